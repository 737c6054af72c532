#!/usr/bin/env python
"""Renders a few waves of the bench scene (for ncu launch lists). Usage: render_once.py [spp] [bounces] [mode]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from cubiquity_b200 import api
spp = int(sys.argv[1]) if len(sys.argv) > 1 else 2
bounces = int(sys.argv[2]) if len(sys.argv) > 2 else 4
mode = int(sys.argv[3]) if len(sys.argv) > 3 else 0
scene = api.Scene("terrain", 12, 1)
ctx = api.Context(0)
ctx.upload(scene.nodes, scene.root, scene.colours)
ctx.set_option("render_mode", mode)
cam = api.default_camera(scene.lower, scene.upper)
acc = torch.zeros(1080 * 1920 * 3, device="cuda")
stream = torch.cuda.current_stream().cuda_stream
ctx.render_device(cam, api.pt_params(1920, 1080, spp=spp, bounces=bounces, variant=1), acc.data_ptr(), stream)
torch.cuda.synchronize()
print("mean", float(acc.mean()) / spp)
