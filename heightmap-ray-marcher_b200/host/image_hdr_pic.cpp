// Radiance HDR (RGBE) and Softimage PIC for the hmap host: the two remaining formats the reference's stbi_load reads
// (vendor/stb_image.h v2.27: stbi__hdr_load :7080-7208 + stbi__hdr_to_ldr :1864-1892, stbi__pic_load :6422-6461).
// Own decoders, written against the file formats and the decoded-pixel contract of those loaders: what matters to the
// ray marcher is the exact 8-bit pixels the reference would have had in memory.
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "image_internal.hpp"

namespace hmrm_host {

namespace {

struct Bytes {
	const std::vector<uint8_t> &f;
	size_t pos;
	explicit Bytes(const std::vector<uint8_t> &file) : f(file), pos(0) {}
	bool eof() const { return pos >= f.size(); }
	int u8() { return pos < f.size() ? f[pos++] : (++pos, 0); }     // past the end: zeros, like stb's get8
	int be16() { const int a = u8(); return (a << 8) | u8(); }
	size_t left() const { return pos < f.size() ? f.size() - pos : 0; }
};

bool fail(std::string *error, const char *why) { *error = why; return false; }

// One header line: the characters up to the newline.  Two quirks of the reference's tokenizer are kept because they
// decide what a header means: a character that is the very last byte of the file is dropped, and a line longer than
// 1022 characters is cut there (the rest of it is skipped).
std::string hdr_line(Bytes &s) {
	std::string line;
	int c = s.u8();
	while (!s.eof() && c != '\n') {
		line.push_back((char)c);
		if (line.size() == 1023) {
			while (!s.eof() && s.u8() != '\n') {}
			break;
		}
		c = s.u8();
	}
	return line;
}

// One RGBE pixel to linear floats (three of them; stb adds a constant alpha of 1 when four channels are requested).
inline void rgbe_to_float(const uint8_t *rgbe, float *out) {
	if (rgbe[3] != 0) {
		const float scale = (float)std::ldexp(1.0f, (int)rgbe[3] - (128 + 8));
		out[0] = rgbe[0] * scale;
		out[1] = rgbe[1] * scale;
		out[2] = rgbe[2] * scale;
	}
	else out[0] = out[1] = out[2] = 0.0f;
}

// Linear float to the 8-bit value stbi_load hands out: gamma 1/2.2, scale 1 (the library defaults; the reference
// never changes them), times 255, plus one half, clamped, truncated.  All in single precision, powf included:
// the reference is compiled as C++, where pow(float, float) is the float overload.
inline uint8_t ldr_byte(float v) {
	static const float inv_gamma = 1.0f / 2.2f;
	float z = std::pow(v * 1.0f, inv_gamma) * 255 + 0.5f;
	if (z < 0) z = 0;
	if (z > 255) z = 255;
	return (uint8_t)(int)z;
}

} // namespace

bool looks_like_hdr(const std::vector<uint8_t> &file) {
	return (file.size() >= 11 && !std::memcmp(file.data(), "#?RADIANCE\n", 11)) ||
	       (file.size() >= 7 && !std::memcmp(file.data(), "#?RGBE\n", 7));
}

bool looks_like_pic(const std::vector<uint8_t> &file) {
	return file.size() >= 92 && !std::memcmp(file.data(), "\x53\x80\xF6\x34", 4) && !std::memcmp(file.data() + 88, "PICT", 4);
}

bool decode_hdr(const std::vector<uint8_t> &file, Decoded *out, std::string *error) {
	Bytes s(file);
	const std::string magic = hdr_line(s);
	if (magic != "#?RADIANCE" && magic != "#?RGBE") return fail(error, "not HDR");
	bool rle_rgbe = false;
	for (;;) {
		const std::string line = hdr_line(s);
		if (line.empty()) break;
		if (line == "FORMAT=32-bit_rle_rgbe") rle_rgbe = true;
	}
	if (!rle_rgbe) return fail(error, "unsupported format");
	const std::string size_line = hdr_line(s);
	const char *t = size_line.c_str();
	if (std::strncmp(t, "-Y ", 3) != 0) return fail(error, "unsupported data layout");
	char *after = NULL;
	const long height = std::strtol(t + 3, &after, 10);
	while (*after == ' ') ++after;
	if (std::strncmp(after, "+X ", 3) != 0) return fail(error, "unsupported data layout");
	const long width = std::strtol(after + 3, NULL, 10);
	if (height > (1 << 24) || width > (1 << 24)) return fail(error, "too large");
	if (height <= 0 || width <= 0 || width * height > (1L << 28)) return fail(error, "bad HDR size");
	const int w = (int)width, h = (int)height;
	const size_t n = (size_t)w * (size_t)h;

	std::vector<float> lin(n * 3);
	// everything from pixel `first` on stored flat, four bytes per pixel
	auto flat_from = [&](size_t first) -> bool {
		if (s.left() < (n - first) * 4) return fail(error, "HDR file too short");
		for (size_t i = first; i < n; ++i, s.pos += 4) rgbe_to_float(&file[s.pos], &lin[i * 3]);
		return true;
	};
	if (w < 8 || w >= 32768) {
		if (!flat_from(0)) return false;
	}
	else {
		std::vector<uint8_t> scan((size_t)w * 4);
		for (int j = 0; j < h; ++j) {
			const int c1 = s.u8(), c2 = s.u8(), hi = s.u8();
			if (c1 != 2 || c2 != 2 || (hi & 0x80)) {
				// Not a run-length scanline.  The reference then takes these four bytes as the FIRST pixel of the image —
				// whichever scanline it is on — and reads all the others flat from here (:7156-7168).
				const uint8_t px[4] = {(uint8_t)c1, (uint8_t)c2, (uint8_t)hi, (uint8_t)s.u8()};
				rgbe_to_float(px, &lin[0]);
				if (!flat_from(1)) return false;
				break;
			}
			if (((hi << 8) | s.u8()) != w) return fail(error, "invalid decoded scanline length");
			for (int k = 0; k < 4; ++k) {
				int i = 0;
				while (i < w) {
					if (s.eof()) return fail(error, "HDR file too short");
					int count = s.u8();
					if (count > 128) {
						const uint8_t value = (uint8_t)s.u8();
						count -= 128;
						if (count > w - i) return fail(error, "bad RLE data in HDR");
						for (int z = 0; z < count; ++z) scan[(size_t)(i++) * 4 + k] = value;
					}
					else {
						if (count > w - i) return fail(error, "bad RLE data in HDR");
						for (int z = 0; z < count; ++z) scan[(size_t)(i++) * 4 + k] = (uint8_t)s.u8();
					}
				}
			}
			for (int i = 0; i < w; ++i) rgbe_to_float(&scan[(size_t)i * 4], &lin[((size_t)j * w + i) * 3]);
		}
	}

	out->w = w;
	out->h = h;
	out->channels = 3;      // four requested: alpha = 1.0 -> 255, which is what convert_channels adds
	out->px.resize(n * 3);
	for (size_t i = 0; i < n * 3; ++i) out->px[i] = ldr_byte(lin[i]);
	return true;
}

bool decode_pic(const std::vector<uint8_t> &file, Decoded *out, std::string *error) {
	Bytes s(file);
	s.pos = 92;
	const int w = s.be16(), h = s.be16();
	if (s.eof()) return fail(error, "file too short (pic header)");
	s.pos += 8;          // ratio, fields, pad
	if (w <= 0 || h <= 0) return fail(error, "bad PIC size");

	struct Packet { int type, channel; };
	Packet packets[10];
	int count = 0, all_channels = 0, chained;
	do {
		if (count == 10) return fail(error, "too many packets");
		chained = s.u8();
		const int size = s.u8();
		packets[count].type = s.u8();
		packets[count].channel = s.u8();
		all_channels |= packets[count].channel;
		++count;
		if (s.eof()) return fail(error, "file too short (reading packets)");
		if (size != 8) return fail(error, "packet isn't 8bpp");
	} while (chained);
	(void)all_channels;       // bit 0x10 = "has alpha" only decides stb's *comp; the pixels are RGBA either way

	// channel mask: 0x80 red, 0x40 green, 0x20 blue, 0x10 alpha; channels no packet carries stay 255
	auto read_value = [&](int channel, uint8_t *dest) -> bool {
		for (int i = 0, mask = 0x80; i < 4; ++i, mask >>= 1)
			if (channel & mask) {
				if (s.eof()) return false;
				dest[i] = (uint8_t)s.u8();
			}
		return true;
	};
	auto copy_value = [](int channel, uint8_t *dest, const uint8_t *src) {
		for (int i = 0, mask = 0x80; i < 4; ++i, mask >>= 1)
			if (channel & mask) dest[i] = src[i];
	};

	out->w = w;
	out->h = h;
	out->channels = 4;
	out->px.assign((size_t)w * (size_t)h * 4, 0xFF);
	for (int y = 0; y < h; ++y) {
		for (int p = 0; p < count; ++p) {
			uint8_t *dest = &out->px[(size_t)y * (size_t)w * 4];
			const int channel = packets[p].channel;
			if (packets[p].type == 0) {           // uncompressed
				for (int x = 0; x < w; ++x, dest += 4)
					if (!read_value(channel, dest)) return fail(error, "PIC file too short");
			}
			else if (packets[p].type == 1) {      // pure run-length: (count, value) pairs, a long count is clipped
				int left = w;
				while (left > 0) {
					int run = s.u8();
					if (s.eof()) return fail(error, "file too short (pure read count)");
					if (run > left) run = left;
					uint8_t value[4];
					if (!read_value(channel, value)) return fail(error, "PIC file too short");
					for (int i = 0; i < run; ++i, dest += 4) copy_value(channel, dest, value);
					left -= run;
				}
			}
			else if (packets[p].type == 2) {      // mixed: >= 128 repeats (128 = 16-bit count follows), < 128 literal run
				int left = w;
				while (left > 0) {
					int run = s.u8();
					if (s.eof()) return fail(error, "file too short (mixed read count)");
					if (run >= 128) {
						run = run == 128 ? s.be16() : run - 127;
						if (run > left) return fail(error, "scanline overrun");
						uint8_t value[4];
						if (!read_value(channel, value)) return fail(error, "PIC file too short");
						for (int i = 0; i < run; ++i, dest += 4) copy_value(channel, dest, value);
					}
					else {
						++run;
						if (run > left) return fail(error, "scanline overrun");
						for (int i = 0; i < run; ++i, dest += 4)
							if (!read_value(channel, dest)) return fail(error, "PIC file too short");
					}
					left -= run;
				}
			}
			else return fail(error, "packet has bad compression type");
		}
	}
	return true;
}

} // namespace hmrm_host
