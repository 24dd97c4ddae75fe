/* TEST INFRASTRUCTURE — not product code.
 *
 * Fake, headless SDL2 surface: just the types, constants and functions that
 * the reference's main/hmap.cpp:10,28,110-112,546-1161 names, so that the
 * UNMODIFIED reference translation unit compiles and runs without a display.
 * Behaviour is scripted through environment variables; see fake_sdl.cpp.
 * SDL2 is not installed in this image and cannot be fetched (no network).
 */
#ifndef HMRM_ORACLE_FAKE_SDL_H
#define HMRM_ORACLE_FAKE_SDL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef uint8_t Uint8;
typedef uint16_t Uint16;
typedef uint32_t Uint32;
typedef int32_t Sint32;

typedef enum { SDL_FALSE = 0, SDL_TRUE = 1 } SDL_bool;

typedef struct SDL_Window SDL_Window;
typedef struct SDL_Renderer SDL_Renderer;
typedef struct SDL_Texture SDL_Texture;

typedef struct SDL_Surface {
	int w, h;
} SDL_Surface;

typedef struct SDL_Color {
	Uint8 r, g, b, a;
} SDL_Color;

typedef struct SDL_Rect {
	int x, y, w, h;
} SDL_Rect;

#define SDL_INIT_VIDEO 0x00000020u
#define SDL_WINDOWPOS_UNDEFINED 0x1FFF0000
#define SDL_WINDOW_RESIZABLE 0x00000020u
#define SDL_WINDOW_FULLSCREEN_DESKTOP 0x00001001u
#define SDL_RENDERER_ACCELERATED 0x00000002u
#define SDL_PIXELFORMAT_ABGR8888 0x16762004u
#define SDL_TEXTUREACCESS_STREAMING 1

enum {
	SDL_QUIT = 0x100,
	SDL_WINDOWEVENT = 0x200,
	SDL_KEYUP = 0x301,
	SDL_TEXTINPUT = 0x303,
	SDL_MOUSEMOTION = 0x400,
	SDL_MOUSEWHEEL = 0x403
};

enum {
	SDL_WINDOWEVENT_SIZE_CHANGED = 6,
	SDL_WINDOWEVENT_FOCUS_GAINED = 12,
	SDL_WINDOWEVENT_FOCUS_LOST = 13
};

typedef Sint32 SDL_Keycode;

enum {
	SDLK_BACKSPACE = '\b',
	SDLK_RETURN = '\r',
	SDLK_1 = '1',
	SDLK_2 = '2',
	SDLK_3 = '3',
	SDLK_BACKQUOTE = '`',
	SDLK_q = 'q',
	SDLK_r = 'r',
	SDLK_F1 = (1 << 30) | 58,
	SDLK_F11 = (1 << 30) | 68,
	SDLK_F12 = (1 << 30) | 69
};

typedef enum {
	KMOD_NONE = 0x0000,
	KMOD_LSHIFT = 0x0001,
	KMOD_RSHIFT = 0x0002,
	KMOD_LCTRL = 0x0040,
	KMOD_RCTRL = 0x0080,
	KMOD_CTRL = KMOD_LCTRL | KMOD_RCTRL,
	KMOD_SHIFT = KMOD_LSHIFT | KMOD_RSHIFT
} SDL_Keymod;

enum {
	SDL_SCANCODE_A = 4,
	SDL_SCANCODE_D = 7,
	SDL_SCANCODE_Q = 20,
	SDL_SCANCODE_S = 22,
	SDL_SCANCODE_W = 26,
	SDL_SCANCODE_SPACE = 44,
	SDL_NUM_SCANCODES = 512
};

typedef struct SDL_Keysym {
	int scancode;
	SDL_Keycode sym;
	Uint16 mod;
} SDL_Keysym;

typedef struct SDL_KeyboardEvent {
	Uint32 type;
	SDL_Keysym keysym;
} SDL_KeyboardEvent;

typedef struct SDL_WindowEvent {
	Uint32 type;
	Uint8 event;
} SDL_WindowEvent;

typedef struct SDL_MouseMotionEvent {
	Uint32 type;
	Sint32 xrel, yrel;
} SDL_MouseMotionEvent;

typedef struct SDL_MouseWheelEvent {
	Uint32 type;
	Sint32 x, y;
} SDL_MouseWheelEvent;

typedef struct SDL_TextInputEvent {
	Uint32 type;
	char text[32];
} SDL_TextInputEvent;

typedef union SDL_Event {
	Uint32 type;
	SDL_KeyboardEvent key;
	SDL_WindowEvent window;
	SDL_MouseMotionEvent motion;
	SDL_MouseWheelEvent wheel;
	SDL_TextInputEvent text;
} SDL_Event;

int SDL_Init(Uint32 flags);
void SDL_Quit(void);
const char *SDL_GetError(void);

SDL_Window *SDL_CreateWindow(const char *title, int x, int y, int w, int h, Uint32 flags);
void SDL_DestroyWindow(SDL_Window *w);
void SDL_SetWindowSize(SDL_Window *w, int width, int height);
void SDL_GetWindowSize(SDL_Window *w, int *width, int *height);
int SDL_SetWindowFullscreen(SDL_Window *w, Uint32 flags);

SDL_Renderer *SDL_CreateRenderer(SDL_Window *w, int index, Uint32 flags);
void SDL_DestroyRenderer(SDL_Renderer *r);
SDL_Texture *SDL_CreateTexture(SDL_Renderer *r, Uint32 format, int access, int w, int h);
SDL_Texture *SDL_CreateTextureFromSurface(SDL_Renderer *r, SDL_Surface *s);
void SDL_DestroyTexture(SDL_Texture *t);
void SDL_FreeSurface(SDL_Surface *s);
int SDL_UpdateTexture(SDL_Texture *t, const SDL_Rect *rect, const void *pixels, int pitch);
int SDL_RenderClear(SDL_Renderer *r);
int SDL_RenderCopy(SDL_Renderer *r, SDL_Texture *t, const SDL_Rect *src, const SDL_Rect *dst);
void SDL_RenderPresent(SDL_Renderer *r);

Uint32 SDL_GetTicks(void);
int SDL_PollEvent(SDL_Event *event);
const Uint8 *SDL_GetKeyboardState(int *numkeys);
SDL_Keymod SDL_GetModState(void);
int SDL_SetRelativeMouseMode(SDL_bool enabled);

#ifdef __cplusplus
}
#endif

#endif
