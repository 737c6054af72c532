/* TEST INFRASTRUCTURE ONLY: a declaration-only stand-in for SDL2's SDL.h.
 *
 * SDL2 is not installed in the build image, and the reference's CPU path tracer
 * (src/application/commands/view/pathtracing_demo.cpp) derives from its SDL window classes. This
 * header declares just enough of SDL's types and functions for the UNMODIFIED pathtracing_demo.cpp,
 * viewer.h and window.h to compile, so that the reference's own bounce loop can be linked into
 * oracle/_ref and called headlessly (oracle/ref_pt_shim.cpp). No window is ever opened; the few SDL
 * functions the reference calls on the paths we use are defined as no-ops in ref_pt_shim.cpp.
 */
#ifndef CBQ_SDL_STUB_H
#define CBQ_SDL_STUB_H

#include <stdint.h>

typedef uint8_t Uint8;
typedef uint16_t Uint16;
typedef uint32_t Uint32;
typedef int32_t Sint32;

typedef enum { SDL_SCANCODE_UNKNOWN = 0, SDL_SCANCODE_A = 4, SDL_SCANCODE_D = 7, SDL_SCANCODE_S = 22, SDL_SCANCODE_W = 26,
               SDL_SCANCODE_LCTRL = 224, SDL_SCANCODE_LSHIFT = 225, SDL_NUM_SCANCODES = 512 } SDL_Scancode;
typedef Sint32 SDL_Keycode;
enum { SDLK_ESCAPE = 27, SDLK_z = 'z', SDLK_F1 = 0x4000003A, SDLK_F2, SDLK_F3, SDLK_F4, SDLK_F5 };

typedef struct SDL_Keysym { SDL_Scancode scancode; SDL_Keycode sym; Uint16 mod; Uint32 unused; } SDL_Keysym;
typedef struct SDL_KeyboardEvent { Uint32 type, timestamp, windowID; Uint8 state, repeat, padding2, padding3; SDL_Keysym keysym; } SDL_KeyboardEvent;
typedef struct SDL_MouseButtonEvent { Uint32 type, timestamp, windowID, which; Uint8 button, state, clicks, padding1; Sint32 x, y; } SDL_MouseButtonEvent;
typedef struct SDL_MouseMotionEvent { Uint32 type, timestamp, windowID, which, state; Sint32 x, y, xrel, yrel; } SDL_MouseMotionEvent;

#define SDL_BUTTON(X) (1 << ((X) - 1))
#define SDL_BUTTON_LEFT 1
#define SDL_BUTTON_MIDDLE 2
#define SDL_BUTTON_RIGHT 3

typedef struct SDL_PixelFormat { Uint32 format; void* palette; Uint8 BitsPerPixel; Uint8 BytesPerPixel; } SDL_PixelFormat;
typedef struct SDL_Rect { int x, y, w, h; } SDL_Rect;
typedef struct SDL_Surface { Uint32 flags; SDL_PixelFormat* format; int w, h; int pitch; void* pixels; } SDL_Surface;
typedef struct SDL_Window SDL_Window;

#ifdef __cplusplus
extern "C" {
#endif
SDL_Surface* SDL_CreateRGBSurface(Uint32 flags, int width, int height, int depth, Uint32 Rmask, Uint32 Gmask, Uint32 Bmask, Uint32 Amask);
void SDL_FreeSurface(SDL_Surface* surface);
int SDL_BlitScaled(SDL_Surface* src, const SDL_Rect* srcrect, SDL_Surface* dst, SDL_Rect* dstrect);
#ifdef __cplusplus
}
#endif

#endif
