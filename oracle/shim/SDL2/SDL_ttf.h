/* TEST INFRASTRUCTURE — not product code. Fake SDL2_ttf surface (main/hmap.cpp:11,561-580,1073-1079). */
#ifndef HMRM_ORACLE_FAKE_SDL_TTF_H
#define HMRM_ORACLE_FAKE_SDL_TTF_H

#include "SDL.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct TTF_Font TTF_Font;

int TTF_Init(void);
void TTF_Quit(void);
const char *TTF_GetError(void);
TTF_Font *TTF_OpenFont(const char *file, int ptsize);
void TTF_CloseFont(TTF_Font *font);
SDL_Surface *TTF_RenderUTF8_Shaded(TTF_Font *font, const char *text, SDL_Color fg, SDL_Color bg);

#ifdef __cplusplus
}
#endif

#endif
