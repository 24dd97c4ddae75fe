// hmap — the reference's binary (`./hmap path/to/config.txt`, main/hmap.cpp:526-544) with a headless
// render-to-PNG mode beside the (SDL) loop.  The config grammar is the reference's; the frame body
// main/hmap.cpp:952-1058 is one call into libhmrm.so (include/hmrm.h).  There is no CPU renderer here.
//
//   hmap config.txt --headless out.png [--projection 1|2|3] [--precision fp64|fp32]
//                   [--traversal auto|brute|skip] [--gpus N] [--stats-json file]
//   hmap config.txt --script frames.txt [--frames N] [--out-prefix P | --record] [--gpus N] ...
//   hmap config.txt --parse-only                    (grammar check: echo + validation, no GPU)
//   hmap --decode-image in.png 3|4 out.raw          (image ingest check, no GPU)
//
// --script: one line of config grammar per frame (what the reference's console accepts, :760-804),
// applied before the frame is rendered — the programmatic animation the reference leaves as a stub
// (:916-926).  Frame n goes to GPU n mod N.  --record names files as the reference's recorder does
// (screenshots/hmap_<id>_<n>.png, :1132-1136) and prints its messages.
// Projection is not part of the grammar (keys 1/2/3 only, :851-868), hence --projection.
// Headless frames are complete images: cycle_period is forced to 1 (the reference's own advice, :913).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <ctime>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include "../../include/hmrm.h"
#include "config.hpp"
#include "image_io.hpp"

using namespace hmrm_host;

#ifdef HMAP_WITH_SDL
namespace hmrm_host {
int run_interactive(Config &cfg, hmrm_ctx *ctx, int precision, int traversal);   // hmap_sdl.cpp
}
#endif

namespace {

struct Options {
	std::string config_path, headless_out, script_path, out_prefix, stats_json;
	int projection, precision, traversal, gpus, frames;
	bool record, parse_only;
	Options() : projection(0), precision(HMRM_FP64_EXACT), traversal(HMRM_TRAVERSAL_AUTO), gpus(1), frames(-1),
	            record(false), parse_only(false) {}
};

void usage_and_exit() {
	std::cerr << "USAGE: hmap.exe path/to/config.txt\n";     // main/hmap.cpp:528
	std::cerr << "       hmap path/to/config.txt --headless out.png [--projection 1|2|3] [--precision fp64|fp32]\n"
	             "            [--traversal auto|brute|skip] [--gpus N] [--script frames.txt [--frames N]\n"
	             "            [--out-prefix P | --record]] [--stats-json file] [--parse-only]\n";
	std::exit(1);
}

struct Device {
	hmrm_ctx *ctx;
	uint8_t *host_frame;      // pinned RGBA8
	size_t host_bytes;
	bool busy;
	int frame_index;          // frame in flight
	hmrm_frame frame;
	Device() : ctx(NULL), host_frame(NULL), host_bytes(0), busy(false), frame_index(-1) {}
};

void die_hmrm(hmrm_ctx *ctx, const char *what) {
	std::cerr << "hmap: " << what << ": " << hmrm_last_error(ctx) << "\n";
	std::exit(1);
}

void push_maps(std::vector<Device> &devs, Config &cfg) {
	for (size_t i = 0; i < devs.size(); ++i) {
		if (cfg.maps_changed &&
		    hmrm_set_maps(devs[i].ctx, cfg.heightmap.pixels.data(), cfg.colormap.pixels.data(), cfg.heightmap.width,
		                  cfg.heightmap.height) != HMRM_OK)
			die_hmrm(devs[i].ctx, "hmrm_set_maps");
		if (cfg.maps_changed || cfg.should_update_heightmap) {
			const double lum[3] = {cfg.lum_r, cfg.lum_g, cfg.lum_b};
			if (hmrm_update_heightmap(devs[i].ctx, lum, cfg.min_height, cfg.max_height) != HMRM_OK)
				die_hmrm(devs[i].ctx, "hmrm_update_heightmap");
		}
	}
	cfg.maps_changed = false;
	cfg.should_update_heightmap = false;
}

hmrm_frame make_frame(const Config &cfg, const Options &opt) {
	hmrm_frame f;
	hmrm_frame_defaults(&f);
	f.projection = opt.projection ? opt.projection : cfg.image_plane;
	f.screen_width = cfg.screen_width;
	f.screen_height = cfg.screen_height;
	f.precision = opt.precision;
	f.cam_pos[0] = cfg.cam_pos[0];
	f.cam_pos[1] = cfg.cam_pos[1];
	f.cam_pos[2] = cfg.cam_pos[2];
	f.hang = cfg.hang;
	f.vang = cfg.vang;
	f.hfov = cfg.hfov;
	f.ortho_width = cfg.ortho_width;
	f.grid_width = cfg.grid_width;
	f.step_dist = cfg.step_dist;
	f.bg[0] = cfg.bg_r;
	f.bg[1] = cfg.bg_g;
	f.bg[2] = cfg.bg_b;
	f.cycle = 0;
	f.cycle_period = 1;
	f.traversal = opt.traversal;
	// counting costs per-warp atomics: only when the caller asked for the numbers (the cut-off status is always reported)
	f.flags = opt.stats_json.empty() ? 0u : HMRM_FLAG_STATS;
	return f;
}

void ensure_host_frame(Device &d, size_t bytes) {
	if (d.host_bytes >= bytes) return;
	if (d.host_frame) hmrm_host_free(d.host_frame);
	void *p = NULL;
	if (hmrm_host_alloc(&p, bytes) != HMRM_OK) {
		std::cerr << "hmap: cannot allocate " << bytes << " bytes of pinned host memory\n";
		std::exit(1);
	}
	d.host_frame = (uint8_t *)p;
	d.host_bytes = bytes;
}

// SavePNG, main/hmap.cpp:157-168
void save_png(const std::vector<uint8_t> &rgba, int w, int h, const std::string &path) {
	std::string why;
	if (!write_png(path, w, h, 4, rgba.data(), &why)) std::cerr << "Failed to write screenshot to " << path << "\n";
	else std::cout << "Saved screenshot at " << path << "\n";
}

struct Totals {
	long long frames, rays, box_hits, surf_hits, steps, fetches;
	double kernel_ms;
	int status;
	Totals() : frames(0), rays(0), box_hits(0), surf_hits(0), steps(0), fetches(0), kernel_ms(0.0), status(0) {}
};

void collect(Device &d, Totals &t) {
	hmrm_stats st;
	if (hmrm_get_stats(d.ctx, &st) != HMRM_OK) die_hmrm(d.ctx, "hmrm_get_stats");
	t.frames += 1;
	t.rays += st.rays;
	t.box_hits += st.box_hits;
	t.surf_hits += st.surf_hits;
	t.steps += st.steps;
	t.fetches += st.fetches;
	t.kernel_ms += st.kernel_ms;
	t.status |= st.status;
}

int decode_image_mode(int argc, char **argv) {
	if (argc != 5) usage_and_exit();
	Image img;
	std::string why;
	if (!load_image(argv[2], std::atoi(argv[3]), &img, &why)) {
		std::cerr << "Failed to load image from " << argv[2] << " (" << why << ")\n";
		return 1;
	}
	std::ofstream out(argv[4], std::ios::binary);
	out.write((const char *)img.pixels.data(), (std::streamsize)img.pixels.size());
	std::cout << img.width << " " << img.height << " " << img.channels << "\n";
	return out ? 0 : 1;
}

} // namespace

int main(int argc, char **argv) {
	if (argc >= 2 && std::strcmp(argv[1], "--decode-image") == 0) return decode_image_mode(argc, argv);
	if (argc < 2) usage_and_exit();

	Options opt;
	opt.config_path = argv[1];
	for (int i = 2; i < argc; ++i) {
		const std::string a = argv[i];
		const bool has_val = i + 1 < argc;
		if (a == "--headless" && has_val) opt.headless_out = argv[++i];
		else if (a == "--projection" && has_val) opt.projection = std::atoi(argv[++i]);
		else if (a == "--precision" && has_val) {
			const std::string v = argv[++i];
			if (v == "fp64") opt.precision = HMRM_FP64_EXACT;
			else if (v == "fp32") opt.precision = HMRM_FP32_FAST;
			else usage_and_exit();
		}
		else if (a == "--traversal" && has_val) {
			const std::string v = argv[++i];
			opt.traversal = v == "brute" ? HMRM_TRAVERSAL_BRUTE : v == "skip" ? HMRM_TRAVERSAL_SKIP : HMRM_TRAVERSAL_AUTO;
		}
		else if (a == "--gpus" && has_val) opt.gpus = std::atoi(argv[++i]);
		else if (a == "--script" && has_val) opt.script_path = argv[++i];
		else if (a == "--frames" && has_val) opt.frames = std::atoi(argv[++i]);
		else if (a == "--out-prefix" && has_val) opt.out_prefix = argv[++i];
		else if (a == "--record") opt.record = true;
		else if (a == "--stats-json" && has_val) opt.stats_json = argv[++i];
		else if (a == "--parse-only") opt.parse_only = true;
		else usage_and_exit();
	}
	if (opt.projection < 0 || opt.projection > 3 || opt.gpus < 1) usage_and_exit();

	// main/hmap.cpp:534-544
	std::ifstream input(opt.config_path.c_str());
	if (!input.is_open()) {
		std::cerr << "Failed to open input file: " << opt.config_path << "\n";
		return 1;
	}
	Config cfg;
	if (consume_config_stream(input, cfg, std::cout, std::cerr) != PARSE_OK) return 1;
	input.close();
	if (opt.parse_only) return 0;

#ifndef HMAP_WITH_SDL
	if (opt.headless_out.empty() && opt.script_path.empty()) {
		std::cerr << "hmap: this build has no SDL window (SDL2 is not available here); use --headless out.png "
		             "or --script frames.txt\n";
		return 1;
	}
#endif

	const int available = hmrm_device_count();
	if (available < 1) {
		std::cerr << "hmap: no CUDA device; there is no CPU renderer\n";
		return 1;
	}
	// HMAP_DEVICE_LIST=2,3,5 picks the CUDA devices --gpus counts over (a device may be listed more than once: several
	// contexts then share it, which is how the multi-device paths are tested on a one-GPU box)
	std::vector<int> device_ids;
	if (const char *list = std::getenv("HMAP_DEVICE_LIST")) {
		std::stringstream ls(list);
		std::string item;
		while (std::getline(ls, item, ',')) {
			const int id = std::atoi(item.c_str());
			if (id < 0 || id >= available) {
				std::cerr << "hmap: HMAP_DEVICE_LIST names device " << id << " but " << available << " device(s) are present\n";
				return 1;
			}
			device_ids.push_back(id);
		}
	}
	else {
		for (int i = 0; i < available; ++i) device_ids.push_back(i);
	}
	if (opt.gpus > (int)device_ids.size()) {
		std::cerr << "hmap: --gpus " << opt.gpus << " but only " << device_ids.size() << " device(s) present\n";
		return 1;
	}
	std::vector<Device> devs((size_t)opt.gpus);
	for (int i = 0; i < opt.gpus; ++i) {
		if (hmrm_create(device_ids[(size_t)i], &devs[(size_t)i].ctx) != HMRM_OK) die_hmrm(NULL, "hmrm_create");
	}
	cfg.maps_changed = true;
	push_maps(devs, cfg);

#ifdef HMAP_WITH_SDL
	if (opt.headless_out.empty() && opt.script_path.empty()) {
		// the reference's interactive mode: `hmap config.txt`
		if (opt.projection) cfg.image_plane = opt.projection;
		const int rc = run_interactive(cfg, devs[0].ctx, opt.precision, opt.traversal);
		for (size_t i = 0; i < devs.size(); ++i) hmrm_destroy(devs[i].ctx);
		return rc;
	}
#endif

	Totals totals;
	const std::clock_t t_start = std::clock();

	if (opt.script_path.empty()) {
		// ---- one frame; with several GPUs its 4-row tile rows are dealt round-robin to the devices (interleaved:
		// sky rows cost nothing and terrain rows do, so contiguous bands balance badly) ----
		hmrm_frame f = make_frame(cfg, opt);
		const size_t bytes = (size_t)f.screen_width * (size_t)f.screen_height * 4;
		ensure_host_frame(devs[0], bytes);
		const int G = opt.gpus;
		for (int g = 0; g < G; ++g) {
			hmrm_frame fb = f;
			if (G > 1) {
				fb.band_count = G;
				fb.band_index = g;
				if (g * 4 >= f.screen_height) continue;       // more devices than tile rows
			}
			// every device copies its own tile rows straight into the one pinned host frame
			if (hmrm_render_async(devs[(size_t)g].ctx, &fb, devs[0].host_frame) != HMRM_OK)
				die_hmrm(devs[(size_t)g].ctx, "hmrm_render");
			devs[(size_t)g].busy = true;
		}
		for (int g = 0; g < G; ++g) {
			if (!devs[(size_t)g].busy) continue;
			if (hmrm_wait(devs[(size_t)g].ctx) != HMRM_OK) die_hmrm(devs[(size_t)g].ctx, "hmrm_wait");
			collect(devs[(size_t)g], totals);
			devs[(size_t)g].busy = false;
		}
		totals.frames = 1;
		std::vector<uint8_t> rgba(devs[0].host_frame, devs[0].host_frame + bytes);
		save_png(rgba, f.screen_width, f.screen_height, opt.headless_out);
	}
	else {
		// ---- recording: one grammar line per frame, frame n on GPU n mod N, PNG encode on worker threads ----
		std::ifstream script(opt.script_path.c_str());
		if (!script.is_open()) {
			std::cerr << "Failed to open input file: " << opt.script_path << "\n";
			return 1;
		}
		std::vector<std::string> lines;
		std::string line;
		while (std::getline(script, line)) {
			if (line.find_first_not_of(" \t\r\n") != std::string::npos) lines.push_back(line);
		}
		int n_frames = opt.frames >= 0 ? opt.frames : (int)lines.size();
		if (opt.record && opt.frames < 0 && lines.empty()) n_frames = cfg.recording_frame_count;
		const std::time_t recording_id = std::time(NULL);      // :879
		std::vector<std::thread> encoders;

		auto frame_path = [&](int n) {
			std::ostringstream ss;
			if (opt.record || opt.out_prefix.empty()) ss << "screenshots/hmap_" << recording_id << "_" << n << ".png";   // :1132-1134
			else ss << opt.out_prefix << n << ".png";
			return ss.str();
		};
		auto retire = [&](Device &d) {
			if (!d.busy) return;
			if (hmrm_wait(d.ctx) != HMRM_OK) die_hmrm(d.ctx, "hmrm_wait");
			collect(d, totals);
			const size_t bytes = (size_t)d.frame.screen_width * (size_t)d.frame.screen_height * 4;
			std::vector<uint8_t> rgba(d.host_frame, d.host_frame + bytes);
			const int w = d.frame.screen_width, h = d.frame.screen_height;
			const std::string path = frame_path(d.frame_index);
			encoders.push_back(std::thread([rgba, w, h, path]() { save_png(rgba, w, h, path); }));
			d.busy = false;
		};

		for (int n = 0; n < n_frames; ++n) {
			if (n < (int)lines.size()) {
				std::istringstream iss(lines[(size_t)n]);
				if (consume_config_stream(iss, cfg, std::cout, std::cerr) != PARSE_OK) return 1;
			}
			Device &d = devs[(size_t)(n % opt.gpus)];
			retire(d);
			if (cfg.maps_changed || cfg.should_update_heightmap) {
				for (size_t i = 0; i < devs.size(); ++i) retire(devs[i]);
				push_maps(devs, cfg);
			}
			d.frame = make_frame(cfg, opt);
			d.frame_index = n;
			ensure_host_frame(d, (size_t)d.frame.screen_width * (size_t)d.frame.screen_height * 4);
			if (hmrm_render_async(d.ctx, &d.frame, d.host_frame) != HMRM_OK) die_hmrm(d.ctx, "hmrm_render");
			d.busy = true;
			if (encoders.size() >= 16) {
				for (size_t i = 0; i < encoders.size(); ++i) encoders[i].join();
				encoders.clear();
			}
		}
		for (size_t i = 0; i < devs.size(); ++i) retire(devs[i]);
		for (size_t i = 0; i < encoders.size(); ++i) encoders[i].join();
		std::cout << "Done recording.\n";      // :1142
	}

	if (totals.status & HMRM_ERR_NONTERMINATING)
		std::cerr << "WARNING: some rays can never leave the grid (the reference would hang); they were cut off\n";

	if (!opt.stats_json.empty()) {
		std::ofstream js(opt.stats_json.c_str());
		js << "{\"frames\": " << totals.frames << ", \"rays\": " << totals.rays << ", \"box_hits\": " << totals.box_hits
		   << ", \"surf_hits\": " << totals.surf_hits << ", \"steps\": " << totals.steps << ", \"fetches\": "
		   << totals.fetches << ", \"kernel_ms\": " << totals.kernel_ms << ", \"status\": " << totals.status
		   << ", \"gpus\": " << opt.gpus << ", \"cpu_seconds\": " << (double)(std::clock() - t_start) / CLOCKS_PER_SEC
		   << "}\n";
	}

	for (size_t i = 0; i < devs.size(); ++i) {
		if (devs[i].host_frame) hmrm_host_free(devs[i].host_frame);
		hmrm_destroy(devs[i].ctx);
	}
	return 0;
}
