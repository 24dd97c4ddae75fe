// TEST INFRASTRUCTURE — not product code.
//
// Headless, scripted stand-in for the ~45 SDL2 / SDL2_ttf entry points that the
// reference's main/hmap.cpp uses (:546-649 init, :687-905 events, :928-929
// keyboard, :1060-1127 present).  Linking the UNMODIFIED reference sources
// against this file gives "Oracle A": the reference's own frame loop
// (main/hmap.cpp:952-1058) rendering into its own framebuf, which this file
// captures at SDL_UpdateTexture (:1082) and times from the last
// SDL_GetModState call (:929) to SDL_UpdateTexture.
//
// Script (environment variables):
//   HMRM_FAKE_PROJ    1|2|3   projection key pressed in the first loop iteration
//                             (main/hmap.cpp:851-868 is the only selector)
//   HMRM_FAKE_FRAMES  N       number of recorded frames (default 1)
//   HMRM_FAKE_WARMUP  K       unrecorded frames rendered first (default 0)
//   HMRM_FAKE_DUMP    prefix  write each recorded frame to <prefix><n>.rgba
//   HMRM_FAKE_TIMES   path    append "<n> <milliseconds>" per recorded frame
//   HMRM_FAKE_SCRIPT  path    one line of config grammar per recorded frame,
//                             typed into the reference's console
//                             (main/hmap.cpp:760-804).  Because look/up are
//                             computed before events are drained (:661-672 vs
//                             :687), every scripted state is rendered twice and
//                             only the second render is recorded.
//   HMRM_FAKE_EVENTS  path    general event script, one event per line "<iteration> <kind> <args>":
//                               key <1|2|3|q|r|backquote|backspace|return|f1|f11|f12> [ctrl] [shift]   (SDL_KEYUP)
//                               text <string>          (SDL_TEXTINPUT)      motion <xrel> <yrel>      wheel <y>
//                               hold <w|a|s|d|q|space> <0|1>  (keyboard state from that iteration on)
//                               resize <w> <h>         (SDL_WINDOWEVENT_SIZE_CHANGED)           quit
//                             every iteration up to HMRM_FAKE_FRAMES is recorded.
#include <SDL2/SDL.h>
#include <SDL2/SDL_ttf.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

namespace {

struct FakeState {
	int proj;
	int frames;
	int warmup;
	std::string dump_prefix;
	std::string times_path;
	std::vector<std::string> script;
	bool scripted;

	int tex_w, tex_h;
	int ticks_calls;
	int filled_iter;
	int recorded;
	std::deque<SDL_Event> queue;
	std::deque<int> queue_mod;          // modifier state while the matching event is being handled
	int current_mod;
	int win_w, win_h;
	struct Scripted {
		int iter;
		std::string kind;
		std::vector<std::string> args;
	};
	std::vector<Scripted> events;
	Uint8 keys[SDL_NUM_SCANCODES];
	std::chrono::steady_clock::time_point t_mod;
	bool init_done;

	FakeState()
		: proj(1), frames(1), warmup(0), scripted(false), tex_w(0), tex_h(0),
		  ticks_calls(0), filled_iter(-1), recorded(0), current_mod(0), win_w(0), win_h(0), init_done(false) {
		std::memset(keys, 0, sizeof keys);
	}
};

FakeState g;

int env_int(const char *name, int dflt) {
	const char *v = std::getenv(name);
	return (v && *v) ? std::atoi(v) : dflt;
}

void lazy_init() {
	if (g.init_done) return;
	g.init_done = true;
	g.proj = env_int("HMRM_FAKE_PROJ", 1);
	g.frames = env_int("HMRM_FAKE_FRAMES", 1);
	g.warmup = env_int("HMRM_FAKE_WARMUP", 0);
	if (const char *v = std::getenv("HMRM_FAKE_DUMP")) g.dump_prefix = v;
	if (const char *v = std::getenv("HMRM_FAKE_TIMES")) g.times_path = v;
	if (const char *v = std::getenv("HMRM_FAKE_SCRIPT")) {
		std::ifstream in(v);
		std::string line;
		while (std::getline(in, line)) {
			if (!line.empty()) g.script.push_back(line);
		}
		g.scripted = true;
		g.frames = (int)g.script.size();
	}
	if (const char *v = std::getenv("HMRM_FAKE_EVENTS")) {
		std::ifstream in(v);
		std::string line;
		while (std::getline(in, line)) {
			std::istringstream ss(line);
			FakeState::Scripted ev;
			if (!(ss >> ev.iter >> ev.kind)) continue;
			if (ev.kind == "text") {
				std::string rest;
				std::getline(ss, rest);
				if (!rest.empty() && rest[0] == ' ') rest.erase(0, 1);
				ev.args.push_back(rest);
			}
			else {
				std::string a;
				while (ss >> a) ev.args.push_back(a);
			}
			g.events.push_back(ev);
		}
	}
}

void push_event(const SDL_Event &e, int mod) {
	g.queue.push_back(e);
	g.queue_mod.push_back(mod);
}

SDL_Keycode key_by_name(const std::string &n) {
	if (n == "1") return SDLK_1;
	if (n == "2") return SDLK_2;
	if (n == "3") return SDLK_3;
	if (n == "q") return SDLK_q;
	if (n == "r") return SDLK_r;
	if (n == "backquote") return SDLK_BACKQUOTE;
	if (n == "backspace") return SDLK_BACKSPACE;
	if (n == "return") return SDLK_RETURN;
	if (n == "f1") return SDLK_F1;
	if (n == "f11") return SDLK_F11;
	if (n == "f12") return SDLK_F12;
	return 0;
}

int scancode_by_name(const std::string &n) {
	if (n == "w") return SDL_SCANCODE_W;
	if (n == "a") return SDL_SCANCODE_A;
	if (n == "s") return SDL_SCANCODE_S;
	if (n == "d") return SDL_SCANCODE_D;
	if (n == "q") return SDL_SCANCODE_Q;
	if (n == "space") return SDL_SCANCODE_SPACE;
	return 0;
}

void fill_scripted_events(int it) {
	for (size_t i = 0; i < g.events.size(); ++i) {
		const FakeState::Scripted &s = g.events[i];
		if (s.iter != it) continue;
		SDL_Event e;
		std::memset(&e, 0, sizeof e);
		if (s.kind == "key" && !s.args.empty()) {
			int mod = KMOD_NONE;
			for (size_t a = 1; a < s.args.size(); ++a) {
				if (s.args[a] == "ctrl") mod |= KMOD_LCTRL;
				if (s.args[a] == "shift") mod |= KMOD_LSHIFT;
			}
			e.type = SDL_KEYUP;
			e.key.keysym.sym = key_by_name(s.args[0]);
			push_event(e, mod);
		}
		else if (s.kind == "text" && !s.args.empty()) {
			for (size_t p = 0; p < s.args[0].size(); p += 31) {
				std::memset(&e, 0, sizeof e);
				e.type = SDL_TEXTINPUT;
				const std::string chunk = s.args[0].substr(p, 31);
				std::memcpy(e.text.text, chunk.c_str(), chunk.size());
				push_event(e, KMOD_NONE);
			}
		}
		else if (s.kind == "motion" && s.args.size() >= 2) {
			e.type = SDL_MOUSEMOTION;
			e.motion.xrel = std::atoi(s.args[0].c_str());
			e.motion.yrel = std::atoi(s.args[1].c_str());
			push_event(e, KMOD_NONE);
		}
		else if (s.kind == "wheel" && !s.args.empty()) {
			e.type = SDL_MOUSEWHEEL;
			e.wheel.y = std::atoi(s.args[0].c_str());
			push_event(e, KMOD_NONE);
		}
		else if (s.kind == "hold" && s.args.size() >= 2) {
			g.keys[scancode_by_name(s.args[0])] = (Uint8)std::atoi(s.args[1].c_str());
		}
		else if (s.kind == "resize" && s.args.size() >= 2) {
			g.win_w = std::atoi(s.args[0].c_str());
			g.win_h = std::atoi(s.args[1].c_str());
			e.type = SDL_WINDOWEVENT;
			e.window.event = SDL_WINDOWEVENT_SIZE_CHANGED;
			push_event(e, KMOD_NONE);
		}
		else if (s.kind == "quit") {
			e.type = SDL_QUIT;
			push_event(e, KMOD_NONE);
		}
	}
}

void push_event(const SDL_Event &e, int mod);

void push_key(SDL_Keycode sym) {
	SDL_Event e;
	std::memset(&e, 0, sizeof e);
	e.type = SDL_KEYUP;
	e.key.keysym.sym = sym;
	push_event(e, KMOD_NONE);
}

void push_text(const std::string &s) {
	for (size_t i = 0; i < s.size(); i += 31) {
		SDL_Event e;
		std::memset(&e, 0, sizeof e);
		e.type = SDL_TEXTINPUT;
		std::string chunk = s.substr(i, 31);
		std::memcpy(e.text.text, chunk.c_str(), chunk.size());
		push_event(e, KMOD_NONE);
	}
}

void fill_scripted_events(int it);

int iteration() { return g.ticks_calls - 2; }

// Which recorded-frame index (or -1) does loop iteration `it` produce?
int recorded_index_of(int it) {
	if (g.scripted) {
		if (it < 0 || (it % 2) == 0) return -1;
		int k = it / 2;
		return k < g.frames ? k : -1;
	}
	int k = it - g.warmup;
	return (k >= 0 && k < g.frames) ? k : -1;
}

int last_iteration() {
	return g.scripted ? 2 * g.frames - 1 : g.warmup + g.frames - 1;
}

void fill_events(int it) {
	if (it == 0) {
		push_key(g.proj == 2 ? SDLK_2 : (g.proj == 3 ? SDLK_3 : SDLK_1));
	}
	if (g.scripted && (it % 2) == 0 && it / 2 < (int)g.script.size()) {
		push_key(SDLK_BACKQUOTE);
		push_text(g.script[(size_t)(it / 2)]);
		push_key(SDLK_RETURN);
	}
	fill_scripted_events(it);
	if (it > last_iteration()) {
		SDL_Event e;
		std::memset(&e, 0, sizeof e);
		e.type = SDL_QUIT;
		push_event(e, KMOD_NONE);
	}
}

int dummy_window, dummy_renderer, dummy_texture, dummy_font;

} // namespace

extern "C" {

int SDL_Init(Uint32) { lazy_init(); return 0; }
void SDL_Quit(void) {}
const char *SDL_GetError(void) { return "fake SDL"; }

SDL_Window *SDL_CreateWindow(const char *, int, int, int, int, Uint32) {
	return (SDL_Window *)&dummy_window;
}
void SDL_DestroyWindow(SDL_Window *) {}
void SDL_SetWindowSize(SDL_Window *, int, int) {}
void SDL_GetWindowSize(SDL_Window *, int *w, int *h) {
	if (w) *w = g.win_w ? g.win_w : g.tex_w;
	if (h) *h = g.win_h ? g.win_h : g.tex_h;
}
int SDL_SetWindowFullscreen(SDL_Window *, Uint32) { return 0; }

SDL_Renderer *SDL_CreateRenderer(SDL_Window *, int, Uint32) {
	return (SDL_Renderer *)&dummy_renderer;
}
void SDL_DestroyRenderer(SDL_Renderer *) {}
SDL_Texture *SDL_CreateTexture(SDL_Renderer *, Uint32, int, int w, int h) {
	g.tex_w = w;
	g.tex_h = h;
	return (SDL_Texture *)&dummy_texture;
}
SDL_Texture *SDL_CreateTextureFromSurface(SDL_Renderer *, SDL_Surface *) {
	return (SDL_Texture *)&dummy_texture;
}
void SDL_DestroyTexture(SDL_Texture *) {}
void SDL_FreeSurface(SDL_Surface *) {}

int SDL_UpdateTexture(SDL_Texture *, const SDL_Rect *, const void *pixels, int pitch) {
	const std::chrono::steady_clock::time_point t1 = std::chrono::steady_clock::now();
	const int k = recorded_index_of(iteration());
	if (k < 0) return 0;

	const double ms = std::chrono::duration<double, std::milli>(t1 - g.t_mod).count();
	if (!g.times_path.empty()) {
		FILE *f = std::fopen(g.times_path.c_str(), "a");
		if (f) {
			std::fprintf(f, "%d %.6f\n", k, ms);
			std::fclose(f);
		}
	}
	if (!g.dump_prefix.empty()) {
		char path[4096];
		std::snprintf(path, sizeof path, "%s%d.rgba", g.dump_prefix.c_str(), k);
		FILE *f = std::fopen(path, "wb");
		if (f) {
			std::fwrite(pixels, 1, (size_t)pitch * (size_t)g.tex_h, f);
			std::fclose(f);
		}
	}
	g.recorded += 1;
	return 0;
}

int SDL_RenderClear(SDL_Renderer *) { return 0; }
int SDL_RenderCopy(SDL_Renderer *, SDL_Texture *, const SDL_Rect *, const SDL_Rect *) { return 0; }

void SDL_RenderPresent(SDL_Renderer *) {
	// All requested frames are captured: leave without rendering a wasted
	// extra frame (quit is only tested at the top of the reference's loop).
	if (iteration() >= last_iteration()) {
		std::fflush(stdout);
		std::exit(0);
	}
}

Uint32 SDL_GetTicks(void) {
	lazy_init();
	g.ticks_calls += 1;
	return (Uint32)(g.ticks_calls * 16);
}

int SDL_PollEvent(SDL_Event *event) {
	const int it = iteration();
	if (it != g.filled_iter) {
		g.filled_iter = it;
		fill_events(it);
	}
	if (g.queue.empty()) {
		g.current_mod = KMOD_NONE;
		return 0;
	}
	if (event) *event = g.queue.front();
	g.current_mod = g.queue_mod.front();
	g.queue.pop_front();
	g.queue_mod.pop_front();
	return 1;
}

const Uint8 *SDL_GetKeyboardState(int *numkeys) {
	if (numkeys) *numkeys = SDL_NUM_SCANCODES;
	return g.keys;
}

SDL_Keymod SDL_GetModState(void) {
	g.t_mod = std::chrono::steady_clock::now();
	return (SDL_Keymod)g.current_mod;
}

int SDL_SetRelativeMouseMode(SDL_bool) { return 0; }

int TTF_Init(void) { return 0; }
void TTF_Quit(void) {}
const char *TTF_GetError(void) { return "fake TTF"; }
TTF_Font *TTF_OpenFont(const char *, int) { return (TTF_Font *)&dummy_font; }
void TTF_CloseFont(TTF_Font *) {}
SDL_Surface *TTF_RenderUTF8_Shaded(TTF_Font *, const char *, SDL_Color, SDL_Color) { return NULL; }

} // extern "C"
