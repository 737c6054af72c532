import sys
rows=[l.rstrip('\n') for l in open(sys.argv[1]) if not l.startswith('total')]
blocks=[]
cur=None
for l in rows:
    parts=l.split(None,4)
    off=parts[0]; ex=int(parts[1]); lanes=parts[2]; smp=int(parts[3]); src=parts[4] if len(parts)>4 else ''
    if cur is None or cur['ex']!=ex or cur['lanes']!=lanes:
        cur={'start':off,'ex':ex,'lanes':lanes,'n':0,'smp':0,'first':src}
        blocks.append(cur)
    cur['n']+=1; cur['smp']+=smp; cur['end']=off; cur['last']=src
tot=sum(b['ex']*b['n'] for b in blocks)
for b in blocks:
    w=b['ex']*b['n']
    if w/tot>0.004:
        print("%s-%s n=%3d ex=%9d lanes=%3s  %5.1f%%  smp=%5d | %s ... %s"%(b['start'],b['end'],b['n'],b['ex'],b['lanes'],100*w/tot,b['smp'],b['first'][:40],b['last'][:40]))
print('total',tot)
