// JPEG ingest: baseline and progressive Huffman JPEG (8 bit, 1 / 3 / 4 components, any integer subsampling ratio,
// restart intervals) decoded to the pixel values stb_image v2.27 returns (vendor/stb_image.h, stbi__jpeg_load with
// req_comp 3 or 4 — what main/hmap.cpp:320-321 and :341-342 call; the reference's own sample_config.txt names
// `path/to/img.jpg`).  An independent implementation of the same arithmetic contract:
//   * inverse DCT: the integer "islow" transform with 12-bit constants, two extra bits kept between the passes,
//     rounding 512 >> 10 and (65536 + (128 << 17)) >> 17 (stb_image.h stbi__idct_block, :2384-2494; its SSE2 path is
//     bit-identical by construction);
//   * chroma upsampling: centred ("jfif") filters (3 near + far + 2) >> 2 vertically / horizontally and
//     (3 t0 + t1 + 8) >> 4 for 2x2, nearest neighbour for every other ratio (:3400-3604);
//   * YCbCr -> RGB in 20.12-style fixed point with the constants rounded to 12 bits and shifted by 8, the Cb term of
//     green masked to its high 16 bits (:3606-3630); Adobe APP14 transform 0 = RGB / CMYK, 2 = YCCK (:3899-3926);
//   * coefficients are 16-bit: products with the quantiser wrap as stb's `short` casts do.
#include "image_internal.hpp"

#include <cstring>

namespace hmrm_host {
namespace {

const uint8_t kZigzag[64 + 15] = {
	0, 1, 8, 16, 9, 2, 3, 10, 17, 24, 32, 25, 18, 11, 4, 5, 12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6, 7, 14, 21, 28,
	35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55,
	62, 63,
	// a corrupt run may index up to 15 places past the end: those land on the last coefficient
	63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63, 63};

struct HuffTable {
	// canonical code description: for every code length 1..16 the first code, the index of its first symbol and the
	// number of codes
	int first_code[17], first_index[17], count[17];
	uint8_t symbol[256];
	int total;
	bool valid;
	HuffTable() : total(0), valid(false) {}
	bool build(const int counts[16]) {
		int code = 0, index = 0;
		for (int len = 1; len <= 16; ++len) {
			first_code[len] = code;
			first_index[len] = index;
			count[len] = counts[len - 1];
			code += counts[len - 1];
			index += counts[len - 1];
			if (counts[len - 1] && code - 1 >= (1 << len)) return false;      // "bad code lengths"
			code <<= 1;
		}
		total = index;
		valid = total <= 256;
		return valid;
	}
};

struct Component {
	int id, h, v, tq, dc_table, ac_table;
	int dc_pred;
	int px_w, px_h;          // effective pixels of this component
	int plane_w, plane_h;    // allocated plane (whole MCUs)
	int blocks_w;            // coefficient blocks per row of the plane
	std::vector<uint8_t> plane;
	std::vector<int16_t> coeff;     // progressive only
};

class JpegDecoder {
public:
	JpegDecoder(const std::vector<uint8_t> &file) : f_(file), pos_(0), bits_(0), nbits_(0), marker_(0), hit_marker_(false) {}

	bool decode(int want_channels, Image *out, std::string *error) {
		if (!run(error)) return false;
		emit(want_channels, out);
		return true;
	}

private:
	const std::vector<uint8_t> &f_;
	size_t pos_;
	// entropy-coded segment reader
	uint32_t bits_;
	int nbits_;
	int marker_;            // marker met while reading entropy-coded data (0 = none)
	bool hit_marker_;

	int width_, height_, ncomp_;
	bool progressive_;
	bool jfif_;
	int adobe_transform_;
	int rgb_ids_;
	int h_max_, v_max_, mcus_x_, mcus_y_;
	int restart_interval_, todo_;
	int scan_n_, order_[4];
	int ss_, se_, ah_, al_;
	int eob_run_;
	Component comp_[4];
	HuffTable dc_[4], ac_[4];
	uint16_t quant_[4][64];

	int get8() { return pos_ < f_.size() ? f_[pos_++] : 0; }
	int get16() { const int a = get8(); return (a << 8) | get8(); }
	bool eof() const { return pos_ >= f_.size(); }

	// ---- marker level ----
	int next_marker() {
		if (marker_) { const int m = marker_; marker_ = 0; return m; }
		int x = get8();
		if (x != 0xFF) return 0;
		while (x == 0xFF) x = get8();
		return x;
	}

	bool fail(std::string *error, const char *why) { *error = std::string("JPEG: ") + why; return false; }

	bool process_marker(int m, std::string *error) {
		switch (m) {
		case 0: return fail(error, "expected marker");
		case 0xDD:
			if (get16() != 4) return fail(error, "bad DRI len");
			restart_interval_ = get16();
			return true;
		case 0xDB: {
			int len = get16() - 2;
			while (len > 0) {
				const int q = get8(), wide = q >> 4, t = q & 15;
				if (wide > 1) return fail(error, "bad DQT type");
				if (t > 3) return fail(error, "bad DQT table");
				for (int i = 0; i < 64; ++i) quant_[t][kZigzag[i]] = (uint16_t)(wide ? get16() : get8());
				len -= wide ? 129 : 65;
			}
			return len == 0 ? true : fail(error, "bad DQT len");
		}
		case 0xC4: {
			int len = get16() - 2;
			while (len > 0) {
				const int q = get8(), cls = q >> 4, id = q & 15;
				if (cls > 1 || id > 3) return fail(error, "bad DHT header");
				int counts[16], n = 0;
				for (int i = 0; i < 16; ++i) { counts[i] = get8(); n += counts[i]; }
				len -= 17;
				HuffTable &t = cls ? ac_[id] : dc_[id];
				if (n > 256 || !t.build(counts)) return fail(error, "bad code lengths");
				for (int i = 0; i < n; ++i) t.symbol[i] = (uint8_t)get8();
				len -= n;
			}
			return len == 0 ? true : fail(error, "bad DHT len");
		}
		default: break;
		}
		if ((m >= 0xE0 && m <= 0xEF) || m == 0xFE) {
			int len = get16();
			if (len < 2) return fail(error, "bad APP/COM len");
			len -= 2;
			if (m == 0xE0 && len >= 5) {
				static const char tag[5] = {'J', 'F', 'I', 'F', 0};
				bool ok = true;
				for (int i = 0; i < 5; ++i) ok = (get8() == (uint8_t)tag[i]) && ok;
				len -= 5;
				if (ok) jfif_ = true;
			}
			else if (m == 0xEE && len >= 12) {
				static const char tag[6] = {'A', 'd', 'o', 'b', 'e', 0};
				bool ok = true;
				for (int i = 0; i < 6; ++i) ok = (get8() == (uint8_t)tag[i]) && ok;
				len -= 6;
				if (ok) {
					get8(); get16(); get16();
					adobe_transform_ = get8();
					len -= 6;
				}
			}
			pos_ += (size_t)len;
			return true;
		}
		return fail(error, "unknown marker");
	}

	bool frame_header(std::string *error) {
		const int len = get16();
		if (len < 11) return fail(error, "bad SOF len");
		if (get8() != 8) return fail(error, "only 8-bit");
		height_ = get16();
		width_ = get16();
		if (height_ == 0) return fail(error, "no header height");
		if (width_ == 0) return fail(error, "0 width");
		ncomp_ = get8();
		if (ncomp_ != 1 && ncomp_ != 3 && ncomp_ != 4) return fail(error, "bad component count");
		if (len != 8 + 3 * ncomp_) return fail(error, "bad SOF len");
		rgb_ids_ = 0;
		h_max_ = v_max_ = 1;
		for (int i = 0; i < ncomp_; ++i) {
			Component &c = comp_[i];
			c.id = get8();
			if (ncomp_ == 3 && c.id == "RGB"[i]) ++rgb_ids_;
			const int q = get8();
			c.h = q >> 4;
			c.v = q & 15;
			if (c.h < 1 || c.h > 4) return fail(error, "bad H");
			if (c.v < 1 || c.v > 4) return fail(error, "bad V");
			c.tq = get8();
			if (c.tq > 3) return fail(error, "bad TQ");
			if (c.h > h_max_) h_max_ = c.h;
			if (c.v > v_max_) v_max_ = c.v;
		}
		for (int i = 0; i < ncomp_; ++i) {
			if (h_max_ % comp_[i].h != 0) return fail(error, "bad H");
			if (v_max_ % comp_[i].v != 0) return fail(error, "bad V");
		}
		if ((long long)width_ * height_ > (1LL << 28)) return fail(error, "too large");
		mcus_x_ = (width_ + h_max_ * 8 - 1) / (h_max_ * 8);
		mcus_y_ = (height_ + v_max_ * 8 - 1) / (v_max_ * 8);
		for (int i = 0; i < ncomp_; ++i) {
			Component &c = comp_[i];
			c.px_w = (width_ * c.h + h_max_ - 1) / h_max_;
			c.px_h = (height_ * c.v + v_max_ - 1) / v_max_;
			c.plane_w = mcus_x_ * c.h * 8;
			c.plane_h = mcus_y_ * c.v * 8;
			c.blocks_w = c.plane_w / 8;
			c.plane.assign((size_t)c.plane_w * (size_t)c.plane_h, 0);
			if (progressive_) c.coeff.assign((size_t)c.plane_w * (size_t)c.plane_h, 0);
			c.dc_table = c.ac_table = 0;
			c.dc_pred = 0;
		}
		return true;
	}

	bool scan_header(std::string *error) {
		const int len = get16();
		scan_n_ = get8();
		if (scan_n_ < 1 || scan_n_ > 4 || scan_n_ > ncomp_) return fail(error, "bad SOS component count");
		if (len != 6 + 2 * scan_n_) return fail(error, "bad SOS len");
		for (int i = 0; i < scan_n_; ++i) {
			const int id = get8(), q = get8();
			int which = 0;
			while (which < ncomp_ && comp_[which].id != id) ++which;
			if (which == ncomp_) return fail(error, "bad SOS component");
			comp_[which].dc_table = q >> 4;
			comp_[which].ac_table = q & 15;
			if (comp_[which].dc_table > 3 || comp_[which].ac_table > 3) return fail(error, "bad huffman table index");
			order_[i] = which;
		}
		ss_ = get8();
		se_ = get8();
		const int a = get8();
		ah_ = a >> 4;
		al_ = a & 15;
		if (progressive_) {
			if (ss_ > 63 || se_ > 63 || ss_ > se_ || ah_ > 13 || al_ > 13) return fail(error, "bad SOS");
		}
		else {
			if (ss_ != 0 || ah_ != 0 || al_ != 0) return fail(error, "bad SOS");
			se_ = 63;
		}
		return true;
	}

	// ---- bit level ----
	void refill() {
		do {
			int b = hit_marker_ ? 0 : get8();
			if (b == 0xFF) {
				int c = get8();
				while (c == 0xFF) c = get8();
				if (c != 0) {
					marker_ = c;
					hit_marker_ = true;
					return;
				}
			}
			bits_ |= (uint32_t)b << (24 - nbits_);
			nbits_ += 8;
		} while (nbits_ <= 24);
	}
	void reset_entropy() {
		bits_ = 0;
		nbits_ = 0;
		hit_marker_ = false;
		marker_ = 0;
		for (int i = 0; i < 4; ++i) comp_[i].dc_pred = 0;
		todo_ = restart_interval_ ? restart_interval_ : 0x7FFFFFFF;
		eob_run_ = 0;
	}
	int take_bits(int n) {                       // n in 0..16, unsigned value
		if (n == 0) return 0;
		if (nbits_ < n) refill();
		const uint32_t v = bits_ >> (32 - n);
		bits_ <<= n;
		nbits_ -= n;
		return (int)v;
	}
	int take_bit() { return take_bits(1); }
	int take_signed(int n) {                     // JPEG "receive and extend"
		if (n == 0) return 0;
		const int v = take_bits(n);
		return v < (1 << (n - 1)) ? v - (1 << n) + 1 : v;
	}
	int decode_symbol(const HuffTable &t) {
		if (!t.valid) return -1;
		if (nbits_ < 16) refill();
		int code = 0;
		for (int len = 1; len <= 16; ++len) {
			code = (int)(bits_ >> (32 - len));
			const int off = code - t.first_code[len];
			if (off >= 0 && off < t.count[len]) {
				if (len > nbits_) return -1;
				bits_ <<= len;
				nbits_ -= len;
				return t.symbol[t.first_index[len] + off];
			}
		}
		nbits_ -= 16;
		return -1;
	}

	// ---- block level ----
	bool baseline_block(int16_t block[64], Component &c, std::string *error) {
		const int t = decode_symbol(dc_[c.dc_table]);
		if (t < 0 || t > 15) return fail(error, "bad huffman code");
		std::memset(block, 0, 64 * sizeof(int16_t));
		const uint16_t *q = quant_[c.tq];
		c.dc_pred += take_signed(t);
		block[0] = (int16_t)(c.dc_pred * q[0]);
		int k = 1;
		do {
			const int rs = decode_symbol(ac_[c.ac_table]);
			if (rs < 0) return fail(error, "bad huffman code");
			const int s = rs & 15, r = rs >> 4;
			if (s == 0) {
				if (rs != 0xF0) break;
				k += 16;
			}
			else {
				k += r;
				const int z = kZigzag[k++];
				block[z] = (int16_t)(take_signed(s) * q[z]);
			}
		} while (k < 64);
		return true;
	}

	bool progressive_dc(int16_t *block, Component &c, std::string *error) {
		if (se_ != 0) return fail(error, "can't merge dc and ac");
		if (ah_ == 0) {
			std::memset(block, 0, 64 * sizeof(int16_t));
			const int t = decode_symbol(dc_[c.dc_table]);
			if (t < 0 || t > 15) return fail(error, "bad huffman code");
			c.dc_pred += take_signed(t);
			block[0] = (int16_t)(c.dc_pred * (1 << al_));
		}
		else if (take_bit()) block[0] = (int16_t)(block[0] + (int16_t)(1 << al_));
		return true;
	}

	void refine(int16_t *p, int16_t bit) {
		if (take_bit() && (*p & bit) == 0) *p = (int16_t)(*p > 0 ? *p + bit : *p - bit);
	}

	bool progressive_ac(int16_t *block, Component &c, std::string *error) {
		if (ss_ == 0) return fail(error, "can't merge dc and ac");
		const HuffTable &table = ac_[c.ac_table];
		if (ah_ == 0) {
			if (eob_run_) { --eob_run_; return true; }
			int k = ss_;
			do {
				const int rs = decode_symbol(table);
				if (rs < 0) return fail(error, "bad huffman code");
				const int s = rs & 15, r = rs >> 4;
				if (s == 0) {
					if (r < 15) {
						eob_run_ = (1 << r);
						if (r) eob_run_ += take_bits(r);
						--eob_run_;
						break;
					}
					k += 16;
				}
				else {
					k += r;
					block[kZigzag[k++]] = (int16_t)(take_signed(s) * (1 << al_));
				}
			} while (k <= se_);
			return true;
		}
		const int16_t bit = (int16_t)(1 << al_);
		if (eob_run_) {
			--eob_run_;
			for (int k = ss_; k <= se_; ++k) {
				int16_t *p = &block[kZigzag[k]];
				if (*p != 0) refine(p, bit);
			}
			return true;
		}
		int k = ss_;
		do {
			const int rs = decode_symbol(table);
			if (rs < 0) return fail(error, "bad huffman code");
			int s = rs & 15, r = rs >> 4;
			if (s == 0) {
				if (r < 15) {
					eob_run_ = (1 << r) - 1;
					if (r) eob_run_ += take_bits(r);
					r = 64;                  // to the end of the band
				}
			}
			else {
				if (s != 1) return fail(error, "bad huffman code");
				s = take_bit() ? bit : -bit;
			}
			while (k <= se_) {
				int16_t *p = &block[kZigzag[k++]];
				if (*p != 0) refine(p, bit);
				else {
					if (r == 0) { *p = (int16_t)s; break; }
					--r;
				}
			}
		} while (k <= se_);
		return true;
	}

	static uint8_t clamp255(int x) { return (uint8_t)(x < 0 ? 0 : (x > 255 ? 255 : x)); }

	// one pass of the 8-point transform on (s0..s7): outputs the even part x0..x3 and the odd part t0..t3
	struct Pass { int x0, x1, x2, x3, t0, t1, t2, t3; };
	static Pass pass(int s0, int s1, int s2, int s3, int s4, int s5, int s6, int s7) {
		static const int c_0_5411961 = (int)(0.5411961f * 4096 + 0.5), c_1_847759065 = (int)(-1.847759065f * 4096 + 0.5),
		                 c_0_765366865 = (int)(0.765366865f * 4096 + 0.5), c_1_175875602 = (int)(1.175875602f * 4096 + 0.5),
		                 c_0_298631336 = (int)(0.298631336f * 4096 + 0.5), c_2_053119869 = (int)(2.053119869f * 4096 + 0.5),
		                 c_3_072711026 = (int)(3.072711026f * 4096 + 0.5), c_1_501321110 = (int)(1.501321110f * 4096 + 0.5),
		                 c_0_899976223 = (int)(-0.899976223f * 4096 + 0.5), c_2_562915447 = (int)(-2.562915447f * 4096 + 0.5),
		                 c_1_961570560 = (int)(-1.961570560f * 4096 + 0.5), c_0_390180644 = (int)(-0.390180644f * 4096 + 0.5);
		Pass r;
		int p1 = (s2 + s6) * c_0_5411961;
		const int e2 = p1 + s6 * c_1_847759065, e3 = p1 + s2 * c_0_765366865;
		const int e0 = (s0 + s4) * 4096, e1 = (s0 - s4) * 4096;
		r.x0 = e0 + e3;
		r.x3 = e0 - e3;
		r.x1 = e1 + e2;
		r.x2 = e1 - e2;
		int t0 = s7, t1 = s5, t2 = s3, t3 = s1;
		int p3 = t0 + t2, p4 = t1 + t3;
		p1 = t0 + t3;
		int p2 = t1 + t2;
		const int p5 = (p3 + p4) * c_1_175875602;
		t0 *= c_0_298631336;
		t1 *= c_2_053119869;
		t2 *= c_3_072711026;
		t3 *= c_1_501321110;
		p1 = p5 + p1 * c_0_899976223;
		p2 = p5 + p2 * c_2_562915447;
		p3 *= c_1_961570560;
		p4 *= c_0_390180644;
		r.t3 = t3 + p1 + p4;
		r.t2 = t2 + p2 + p3;
		r.t1 = t1 + p2 + p4;
		r.t0 = t0 + p1 + p3;
		return r;
	}

	static void idct(uint8_t *out, int stride, const int16_t d[64]) {
		int mid[64];
		for (int i = 0; i < 8; ++i) {
			Pass p = pass(d[i], d[8 + i], d[16 + i], d[24 + i], d[32 + i], d[40 + i], d[48 + i], d[56 + i]);
			p.x0 += 512; p.x1 += 512; p.x2 += 512; p.x3 += 512;
			mid[i] = (p.x0 + p.t3) >> 10;
			mid[56 + i] = (p.x0 - p.t3) >> 10;
			mid[8 + i] = (p.x1 + p.t2) >> 10;
			mid[48 + i] = (p.x1 - p.t2) >> 10;
			mid[16 + i] = (p.x2 + p.t1) >> 10;
			mid[40 + i] = (p.x2 - p.t1) >> 10;
			mid[24 + i] = (p.x3 + p.t0) >> 10;
			mid[32 + i] = (p.x3 - p.t0) >> 10;
		}
		for (int i = 0; i < 8; ++i) {
			const int *v = mid + 8 * i;
			uint8_t *o = out + (size_t)i * stride;
			Pass p = pass(v[0], v[1], v[2], v[3], v[4], v[5], v[6], v[7]);
			const int bias = 65536 + (128 << 17);
			p.x0 += bias; p.x1 += bias; p.x2 += bias; p.x3 += bias;
			o[0] = clamp255((p.x0 + p.t3) >> 17);
			o[7] = clamp255((p.x0 - p.t3) >> 17);
			o[1] = clamp255((p.x1 + p.t2) >> 17);
			o[6] = clamp255((p.x1 - p.t2) >> 17);
			o[2] = clamp255((p.x2 + p.t1) >> 17);
			o[5] = clamp255((p.x2 - p.t1) >> 17);
			o[3] = clamp255((p.x3 + p.t0) >> 17);
			o[4] = clamp255((p.x3 - p.t0) >> 17);
		}
	}

	// after every MCU: count down the restart interval; false = stop the scan (no restart marker where one is due)
	bool mcu_done() {
		if (--todo_ > 0) return true;
		if (nbits_ < 24) refill();
		if (!(marker_ >= 0xD0 && marker_ <= 0xD7)) return false;
		reset_entropy();
		return true;
	}

	bool scan(std::string *error) {
		reset_entropy();
		int16_t block[64];
		if (scan_n_ == 1) {
			Component &c = comp_[order_[0]];
			const int bw = (c.px_w + 7) >> 3, bh = (c.px_h + 7) >> 3;
			for (int j = 0; j < bh; ++j) {
				for (int i = 0; i < bw; ++i) {
					if (!progressive_) {
						if (!baseline_block(block, c, error)) return false;
						idct(&c.plane[(size_t)c.plane_w * j * 8 + (size_t)i * 8], c.plane_w, block);
					}
					else {
						int16_t *b = &c.coeff[64 * ((size_t)i + (size_t)j * c.blocks_w)];
						if (!(ss_ == 0 ? progressive_dc(b, c, error) : progressive_ac(b, c, error))) return false;
					}
					if (!mcu_done()) return true;
				}
			}
			return true;
		}
		for (int j = 0; j < mcus_y_; ++j) {
			for (int i = 0; i < mcus_x_; ++i) {
				for (int k = 0; k < scan_n_; ++k) {
					Component &c = comp_[order_[k]];
					for (int y = 0; y < c.v; ++y) {
						for (int x = 0; x < c.h; ++x) {
							const int bx = i * c.h + x, by = j * c.v + y;
							if (!progressive_) {
								if (!baseline_block(block, c, error)) return false;
								idct(&c.plane[(size_t)c.plane_w * by * 8 + (size_t)bx * 8], c.plane_w, block);
							}
							else if (!progressive_dc(&c.coeff[64 * ((size_t)bx + (size_t)by * c.blocks_w)], c, error)) return false;
						}
					}
				}
				if (!mcu_done()) return true;
			}
		}
		return true;
	}

	void finish_progressive() {
		for (int n = 0; n < ncomp_; ++n) {
			Component &c = comp_[n];
			const int bw = (c.px_w + 7) >> 3, bh = (c.px_h + 7) >> 3;
			for (int j = 0; j < bh; ++j) {
				for (int i = 0; i < bw; ++i) {
					int16_t *b = &c.coeff[64 * ((size_t)i + (size_t)j * c.blocks_w)];
					for (int k = 0; k < 64; ++k) b[k] = (int16_t)(b[k] * quant_[c.tq][k]);
					idct(&c.plane[(size_t)c.plane_w * j * 8 + (size_t)i * 8], c.plane_w, b);
				}
			}
		}
	}

	bool run(std::string *error) {
		jfif_ = false;
		adobe_transform_ = -1;
		restart_interval_ = 0;
		std::memset(quant_, 0, sizeof quant_);
		if (next_marker() != 0xD8) return fail(error, "no SOI");
		int m = next_marker();
		while (!(m == 0xC0 || m == 0xC1 || m == 0xC2)) {
			if (!process_marker(m, error)) return false;
			m = next_marker();
			while (m == 0) {
				if (eof()) return fail(error, "no SOF");
				m = next_marker();
			}
		}
		progressive_ = m == 0xC2;
		if (!frame_header(error)) return false;
		m = next_marker();
		while (m != 0xD9) {
			if (m == 0xDA) {
				if (!scan_header(error) || !scan(error)) return false;
				if (marker_ == 0) {
					// zero padding after the entropy-coded data: look for the next marker
					while (!eof()) {
						if (get8() == 0xFF) { marker_ = get8(); break; }
					}
				}
			}
			else if (m == 0xDC) {
				if (get16() != 4) return fail(error, "bad DNL len");
				if (get16() != height_) return fail(error, "bad DNL height");
			}
			else if (!process_marker(m, error)) return false;
			m = next_marker();
		}
		if (progressive_) finish_progressive();
		return true;
	}

	// ---- upsampling + colour ----
	static uint8_t blinn(int x, int y) {
		const unsigned t = (unsigned)(x * y + 128);
		return (uint8_t)((t + (t >> 8)) >> 8);
	}

	void emit(int n, Image *out) const {
		out->width = width_;
		out->height = height_;
		out->channels = n;
		out->pixels.assign((size_t)width_ * (size_t)height_ * (size_t)n, 0);
		const bool is_rgb = ncomp_ == 3 && (rgb_ids_ == 3 || (adobe_transform_ == 0 && !jfif_));
		struct Up { int hs, vs, ystep, ypos, w_lo; size_t line0, line1; };
		Up up[4];
		std::vector<uint8_t> row[4];
		for (int k = 0; k < ncomp_; ++k) {
			up[k].hs = h_max_ / comp_[k].h;
			up[k].vs = v_max_ / comp_[k].v;
			up[k].ystep = up[k].vs >> 1;
			up[k].w_lo = (width_ + up[k].hs - 1) / up[k].hs;
			up[k].ypos = 0;
			up[k].line0 = up[k].line1 = 0;
			row[k].assign((size_t)width_ + 8, 0);
		}
		for (int j = 0; j < height_; ++j) {
			const uint8_t *line[4] = {NULL, NULL, NULL, NULL};
			for (int k = 0; k < ncomp_; ++k) {
				Up &u = up[k];
				const bool bottom = u.ystep >= (u.vs >> 1);
				const uint8_t *near_ = &comp_[k].plane[bottom ? u.line1 : u.line0];
				const uint8_t *far_ = &comp_[k].plane[bottom ? u.line0 : u.line1];
				uint8_t *o = row[k].data();
				const int w = u.w_lo;
				if (u.hs == 1 && u.vs == 1) line[k] = near_;
				else {
					if (u.hs == 1 && u.vs == 2) {
						for (int i = 0; i < w; ++i) o[i] = (uint8_t)((3 * near_[i] + far_[i] + 2) >> 2);
					}
					else if (u.hs == 2 && u.vs == 1) {
						if (w == 1) o[0] = o[1] = near_[0];
						else {
							o[0] = near_[0];
							o[1] = (uint8_t)((near_[0] * 3 + near_[1] + 2) >> 2);
							int i;
							for (i = 1; i < w - 1; ++i) {
								const int c3 = 3 * near_[i] + 2;
								o[i * 2] = (uint8_t)((c3 + near_[i - 1]) >> 2);
								o[i * 2 + 1] = (uint8_t)((c3 + near_[i + 1]) >> 2);
							}
							o[i * 2] = (uint8_t)((near_[w - 2] * 3 + near_[w - 1] + 2) >> 2);
							o[i * 2 + 1] = near_[w - 1];
						}
					}
					else if (u.hs == 2 && u.vs == 2) {
						if (w == 1) o[0] = o[1] = (uint8_t)((3 * near_[0] + far_[0] + 2) >> 2);
						else {
							int t1 = 3 * near_[0] + far_[0];
							o[0] = (uint8_t)((t1 + 2) >> 2);
							for (int i = 1; i < w; ++i) {
								const int t0 = t1;
								t1 = 3 * near_[i] + far_[i];
								o[i * 2 - 1] = (uint8_t)((3 * t0 + t1 + 8) >> 4);
								o[i * 2] = (uint8_t)((3 * t1 + t0 + 8) >> 4);
							}
							o[w * 2 - 1] = (uint8_t)((t1 + 2) >> 2);
						}
					}
					else {
						for (int i = 0; i < w; ++i) {
							for (int r = 0; r < u.hs; ++r) {
								if ((size_t)(i * u.hs + r) < row[k].size()) o[i * u.hs + r] = near_[i];
							}
						}
					}
					line[k] = o;
				}
				if (++u.ystep >= u.vs) {
					u.ystep = 0;
					u.line0 = u.line1;
					if (++u.ypos < comp_[k].px_h) u.line1 += (size_t)comp_[k].plane_w;
				}
			}
			uint8_t *o = &out->pixels[(size_t)j * (size_t)width_ * (size_t)n];
			for (int i = 0; i < width_; ++i, o += n) {
				int r, g, b;
				if (ncomp_ == 1) r = g = b = line[0][i];
				else if (ncomp_ == 3 && is_rgb) { r = line[0][i]; g = line[1][i]; b = line[2][i]; }
				else if (ncomp_ == 4 && adobe_transform_ == 0) {
					const int m = line[3][i];
					r = blinn(line[0][i], m); g = blinn(line[1][i], m); b = blinn(line[2][i], m);
				}
				else {
					const int yf = (line[0][i] << 20) + (1 << 19);
					const int cr = line[2][i] - 128, cb = line[1][i] - 128;
					const int k_r = ((int)(1.40200f * 4096.0f + 0.5f)) << 8, k_g1 = ((int)(0.71414f * 4096.0f + 0.5f)) << 8;
					const int k_g2 = ((int)(0.34414f * 4096.0f + 0.5f)) << 8, k_b = ((int)(1.77200f * 4096.0f + 0.5f)) << 8;
					r = (yf + cr * k_r) >> 20;
					g = (yf + cr * -k_g1 + (int)((unsigned)(cb * -k_g2) & 0xFFFF0000u)) >> 20;
					b = (yf + cb * k_b) >> 20;
					r = clamp255(r); g = clamp255(g); b = clamp255(b);
					if (ncomp_ == 4 && adobe_transform_ == 2) {
						const int m = line[3][i];
						r = blinn(255 - r, m); g = blinn(255 - g, m); b = blinn(255 - b, m);
					}
				}
				o[0] = (uint8_t)r; o[1] = (uint8_t)g; o[2] = (uint8_t)b;
				if (n == 4) o[3] = 255;
			}
		}
	}
};

} // namespace

bool decode_jpeg(const std::vector<uint8_t> &file, int want_channels, Image *out, std::string *error) {
	JpegDecoder d(file);
	return d.decode(want_channels, out, error);
}

} // namespace hmrm_host
