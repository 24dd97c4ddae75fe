// TEST INFRASTRUCTURE — not product code.
//
// Include-harness around the UNMODIFIED reference translation unit: the
// reference's main() is renamed by the preprocessor and main/hmap.cpp is
// included where it lies under /root/reference (nothing is copied), so that
// its file-static functions and globals (UpdateHeightmap main/hmap.cpp:171,
// DegreesToRads :131, heightmap_buf :53 ...) and the classes in src/*.cpp can
// be driven directly to produce known-answer vectors (bit patterns of
// intermediates).  Built into oracle/_ref/ref_harness by oracle/Makefile.
//
//   ref_harness heights <rgb8.raw> W H lr lg lb min max <out.f64>
//   ref_harness kat < queries > answers       (hex-float text protocol)
//   ref_harness decode <image> 3|4 <out.raw>  (stbi_load exactly as main/hmap.cpp:320-321 / :341-342 call it)
//
// kat queries (all numbers C99 hex floats, "%la"):
//   R proj px py pz hang vang hfov ow W H w h   -> ray through the reference's own
//       camera-basis expressions (restated from :661-672, the only lines of
//       main() that cannot be called) + the reference ImagePlane classes
//   D ox oy oz dx dy dz c0x c0y c0z c1x c1y c1z -> distance(), intersection()
//   G deg                                        -> DegreesToRads(deg)
#define main hmap_reference_main
#include "main/hmap.cpp"
#undef main

#include <cstdio>
#include <cstring>
#include <vector>

static int cmd_heights(int argc, char **argv) {
	if (argc != 11) {
		std::fprintf(stderr, "usage: heights rgb.raw W H lr lg lb min max out.f64\n");
		return 2;
	}
	heightmap_width = std::atoi(argv[3]);
	heightmap_height = std::atoi(argv[4]);
	lum_r = std::strtod(argv[5], NULL);
	lum_g = std::strtod(argv[6], NULL);
	lum_b = std::strtod(argv[7], NULL);
	min_height = std::strtod(argv[8], NULL);
	max_height = std::strtod(argv[9], NULL);

	const size_t n = (size_t)heightmap_width * (size_t)heightmap_height;
	std::vector<unsigned char> rgb(n * 3);
	FILE *f = std::fopen(argv[2], "rb");
	if (!f || std::fread(&rgb[0], 1, n * 3, f) != n * 3) {
		std::fprintf(stderr, "cannot read %s\n", argv[2]);
		return 1;
	}
	std::fclose(f);

	base_heightmap_buf = &rgb[0];
	UpdateHeightmap();

	f = std::fopen(argv[10], "wb");
	if (!f || std::fwrite(heightmap_buf, sizeof(double), n, f) != n) {
		std::fprintf(stderr, "cannot write %s\n", argv[10]);
		return 1;
	}
	std::fclose(f);
	return 0;
}

static int cmd_kat() {
	char tag[8];
	while (std::scanf("%7s", tag) == 1) {
		if (tag[0] == 'R') {
			int proj, W, H;
			double px, py, pz, ha, va, hf, ow, w, h;
			if (std::scanf("%d %la %la %la %la %la %la %la %d %d %la %la",
					&proj, &px, &py, &pz, &ha, &va, &hf, &ow, &W, &H, &w, &h) != 12) return 1;

			// main/hmap.cpp:661-672, restated (inline in main(), not callable)
			glm::dvec3 look(sin(va) * cos(ha), sin(va) * sin(ha), cos(va));
			double up_vang = va - (M_PI / 2.0);
			glm::dvec3 up(sin(up_vang) * cos(ha), sin(up_vang) * sin(ha), cos(up_vang));
			glm::dvec3 pos(px, py, pz);

			ImagePlane *ip;
			if (proj == IMAGEPLANE_PERSPECTIVE) ip = new Perspective(pos, look, up, hf, (double)W / H);
			else if (proj == IMAGEPLANE_SPHERICAL) ip = new Spherical(pos, ha, va, hf, (double)W / H);
			else ip = new Orthographic(pos, look, up, ow, W, H);

			struct Ray r = ip->GetRay(w, h);
			std::printf("R %a %a %a %a %a %a\n", r.pos.x, r.pos.y, r.pos.z, r.dir.x, r.dir.y, r.dir.z);
			delete ip;
		}
		else if (tag[0] == 'D') {
			double v[12];
			for (int i = 0; i < 12; ++i) {
				if (std::scanf("%la", &v[i]) != 1) return 1;
			}
			struct Ray r = {glm::dvec3(v[0], v[1], v[2]), glm::dvec3(v[3], v[4], v[5])};
			glm::dvec3 c0(v[6], v[7], v[8]), c1(v[9], v[10], v[11]);
			double d = distance(r, c0, c1);
			glm::dvec3 out(0.0, 0.0, 0.0);
			bool hit = intersection(&out, r, c0, c1);
			std::printf("D %a %d %a %a %a\n", d, hit ? 1 : 0, out.x, out.y, out.z);
		}
		else if (tag[0] == 'G') {
			double deg;
			if (std::scanf("%la", &deg) != 1) return 1;
			std::printf("G %a\n", DegreesToRads(deg));
		}
		else {
			return 1;
		}
	}
	return 0;
}

static int cmd_decode(int argc, char **argv) {
	if (argc != 5) return 2;
	int w = 0, h = 0, n = 0;
	const int comp = std::atoi(argv[3]);
	unsigned char *px = stbi_load(argv[2], &w, &h, &n, comp);
	if (px == NULL) {
		std::printf("FAIL\n");
		return 0;
	}
	FILE *f = std::fopen(argv[4], "wb");
	std::fwrite(px, 1, (size_t)w * (size_t)h * (size_t)comp, f);
	std::fclose(f);
	std::printf("%d %d %d\n", w, h, comp);
	stbi_image_free(px);
	return 0;
}

int main(int argc, char **argv) {
	if (argc >= 2 && std::strcmp(argv[1], "decode") == 0) return cmd_decode(argc, argv);
	if (argc >= 2 && std::strcmp(argv[1], "heights") == 0) return cmd_heights(argc, argv);
	if (argc >= 2 && std::strcmp(argv[1], "kat") == 0) return cmd_kat();
	std::fprintf(stderr, "usage: ref_harness heights ...|kat\n");
	return 2;
}
