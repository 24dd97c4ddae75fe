// Robustness driver for the host's image decoders (tests/test_image_fuzz.py builds it with -fsanitize=address,undefined):
// every file given is decoded as it is, then truncated at many lengths and with bytes overwritten at pseudo-random
// places.  A damaged file may decode or be refused; it must never crash, read out of bounds, or run away.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "image_io.hpp"

static unsigned long long rng_state = 0x9E3779B97F4A7C15ULL;
static unsigned rnd() {
	rng_state = rng_state * 6364136223846793005ULL + 1442695040888963407ULL;
	return (unsigned)(rng_state >> 33);
}

int main(int argc, char **argv) {
	const int variants = argc > 1 ? std::atoi(argv[1]) : 200;
	long decoded = 0, refused = 0;
	for (int a = 2; a < argc; ++a) {
		std::vector<uint8_t> file;
		FILE *f = std::fopen(argv[a], "rb");
		if (!f) { std::fprintf(stderr, "cannot open %s\n", argv[a]); return 2; }
		uint8_t buf[65536];
		size_t n;
		while ((n = std::fread(buf, 1, sizeof buf, f)) > 0) file.insert(file.end(), buf, buf + n);
		std::fclose(f);
		hmrm_host::Image img;
		std::string err;
		for (int comp = 3; comp <= 4; ++comp)
			if (!hmrm_host::load_image_memory(file, comp, &img, &err)) {
				// the undamaged fixture: only the one stb itself cannot decode may be refused
				if (std::string(argv[a]).find("rgb16.ppm") == std::string::npos) {
					std::fprintf(stderr, "%s refused undamaged: %s\n", argv[a], err.c_str());
					return 3;
				}
			}
		for (int v = 0; v < variants; ++v) {
			std::vector<uint8_t> bad(file);
			const unsigned kind = rnd() % 3;
			if (kind == 0 || kind == 2) bad.resize(file.size() > 1 ? 1 + rnd() % (file.size() - 1) : 0);
			if ((kind == 1 || kind == 2) && !bad.empty()) {
				const int flips = 1 + (int)(rnd() % 4);
				for (int i = 0; i < flips; ++i) {
					// mostly in the headers, where the sizes and table definitions live
					const size_t span = (rnd() % 4) ? (bad.size() < 128 ? bad.size() : 128) : bad.size();
					bad[rnd() % span] = (uint8_t)rnd();
				}
			}
			hmrm_host::Image out;
			if (hmrm_host::load_image_memory(bad, 3 + (int)(rnd() % 2), &out, &err)) {
				decoded += 1;
				if (out.pixels.size() != (size_t)out.width * (size_t)out.height * (size_t)out.channels) {
					std::fprintf(stderr, "%s variant %d: inconsistent image\n", argv[a], v);
					return 4;
				}
			}
			else refused += 1;
		}
	}
	std::printf("%ld decoded, %ld refused\n", decoded, refused);
	return 0;
}
