// Host side of the boundary: per-frame constants of the image plane.
//
// The reference constructs one ImagePlane subclass per frame on the host
// (main/hmap.cpp:952-965) from look/up (:661-672); the per-pixel GetRay then
// only needs a handful of vectors.  This header does the constructor maths on
// the host in FP64 with the same libm (sin/cos/tan) and the same operation
// order, so the device kernel can stay free of transcendental functions:
//   Perspective  src/Perspective.cpp:3-23    -> ul, pr, pd, cam
//   Spherical    src/Spherical.cpp:3-15      -> per-column cos/sin(ha), per-row sin/cos(va) tables
//                                               (ha depends on the column only, va on the row only, :18-19)
//   Orthographic src/Orthographic.cpp:3-17   -> ul, pr, pd, float-rounded look
// Compiled with -ffp-contract=off: every operation below is a single rounding.
#ifndef HMRM_FRAME_SETUP_H
#define HMRM_FRAME_SETUP_H

#include <cmath>
#include <vector>

#include "../../include/hmrm.h"

namespace hmrm {

struct Vec3 {
	double x, y, z;
};

inline Vec3 vmake(double x, double y, double z) { Vec3 r = {x, y, z}; return r; }
inline Vec3 vadd(Vec3 a, Vec3 b) { return vmake(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vec3 vsub(Vec3 a, Vec3 b) { return vmake(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vec3 vneg(Vec3 a) { return vmake(-a.x, -a.y, -a.z); }
inline Vec3 vscale(double s, Vec3 a) { return vmake(s * a.x, s * a.y, s * a.z); }
// GLM 0.9.9.8 scalar cross / normalize (see DESIGN.md "third-party arithmetic")
inline Vec3 vcross(Vec3 a, Vec3 b) {
	return vmake(a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y);
}
inline Vec3 vnormalize(Vec3 a) {
	const double d = (a.x * a.x + a.y * a.y) + a.z * a.z;
	const double inv = 1.0 / std::sqrt(d);
	return vmake(a.x * inv, a.y * inv, a.z * inv);
}

// main/hmap.cpp:131-133
inline double deg2rad(double degrees) { return (degrees / 180.0) * M_PI; }

// main/hmap.cpp:661-672
inline void camera_basis(double hang, double vang, Vec3 *look, Vec3 *up) {
	const double sv = std::sin(vang), cv = std::cos(vang);
	const double sh = std::sin(hang), ch = std::cos(hang);
	*look = vmake(sv * ch, sv * sh, cv);
	const double uv = vang - (M_PI / 2.0);
	const double su = std::sin(uv), cu = std::cos(uv);
	*up = vmake(su * ch, su * sh, cu);
}

struct PlaneConst {
	int projection;
	Vec3 cam;
	Vec3 ul, pr, pd;     // upper_left, plane_right, plane_down (perspective, orthographic)
	Vec3 look;           // orthographic ray direction (float-rounded look)
	double hfov, vfov, ul_hang, ul_vang;   // spherical
};

inline PlaneConst build_plane(const hmrm_frame &f) {
	PlaneConst pc;
	const int W = f.screen_width, H = f.screen_height;
	const double ar = (double)W / H;   // main/hmap.cpp:956,960
	Vec3 look, up;
	camera_basis(f.hang, f.vang, &look, &up);
	pc.projection = f.projection;
	pc.cam = vmake(f.cam_pos[0], f.cam_pos[1], f.cam_pos[2]);
	pc.ul = pc.pr = pc.pd = pc.look = vmake(0.0, 0.0, 0.0);
	pc.hfov = pc.vfov = pc.ul_hang = pc.ul_vang = 0.0;

	if (f.projection == HMRM_PERSPECTIVE) {
		const double half_w = std::tan(f.hfov / 2.0);
		const double half_h = half_w / ar;
		const Vec3 right = vnormalize(vcross(look, up));
		const Vec3 centre = vadd(pc.cam, look);
		const Vec3 to_top = vscale(half_h, up);
		const Vec3 to_right = vscale(half_w, right);
		pc.ul = vsub(vadd(centre, to_top), to_right);
		const Vec3 lower_left = vsub(vsub(centre, to_top), to_right);
		const Vec3 upper_right = vadd(vadd(centre, to_top), to_right);
		pc.pr = vsub(upper_right, pc.ul);
		pc.pd = vsub(lower_left, pc.ul);
	}
	else if (f.projection == HMRM_SPHERICAL) {
		pc.hfov = f.hfov;
		pc.vfov = f.hfov / ar;
		pc.ul_hang = f.hang + (f.hfov / 2.0);
		pc.ul_vang = f.vang - (pc.vfov / 2.0);
	}
	else {
		// look and up are narrowed to float by the constructor's parameter types
		const Vec3 lf = vmake((double)(float)look.x, (double)(float)look.y, (double)(float)look.z);
		const Vec3 uf = vmake((double)(float)up.x, (double)(float)up.y, (double)(float)up.z);
		const Vec3 right = vcross(lf, uf);   // not normalised
		const double half_w = (W / 2.0) * f.ortho_width;
		const double half_h = (H / 2.0) * f.ortho_width;
		pc.look = lf;
		pc.ul = vadd(vsub(pc.cam, vscale(half_w, right)), vscale(half_h, uf));
		pc.pr = vscale(W * f.ortho_width, right);
		pc.pd = vscale(H * f.ortho_width, vneg(uf));
	}
	return pc;
}

// Host mirror of ImagePlane::GetRay (src/Perspective.cpp:25-32, src/Spherical.cpp:17-31,
// src/Orthographic.cpp:19-25).  Not used by the render path; exported as hmrm_get_ray.
inline void plane_ray(const PlaneConst &pc, double w, double h, Vec3 *pos, Vec3 *dir) {
	if (pc.projection == HMRM_PERSPECTIVE) {
		const Vec3 q = vadd(vadd(pc.ul, vscale(w, pc.pr)), vscale(h, pc.pd));
		*pos = pc.cam;
		*dir = vnormalize(vsub(q, pc.cam));
	}
	else if (pc.projection == HMRM_SPHERICAL) {
		const double ha = pc.ul_hang - w * pc.hfov;
		const double va = pc.ul_vang + h * pc.vfov;
		const double sv = std::sin(va);
		*pos = pc.cam;
		*dir = vmake(sv * std::cos(ha), sv * std::sin(ha), std::cos(va));
	}
	else {
		*pos = vadd(vadd(pc.ul, vscale(w, pc.pr)), vscale(h, pc.pd));
		*dir = pc.look;
	}
}

// w = px/(W-1) per column and h = py/(H-1) per row (main/hmap.cpp:985-988): true divisions,
// done once per resolution on the host instead of twice per pixel on the device.
inline void fill_wh_tables(int W, int H, std::vector<double> *wtab, std::vector<double> *htab) {
	wtab->resize((size_t)W);
	htab->resize((size_t)H);
	for (int x = 0; x < W; ++x) (*wtab)[(size_t)x] = (double)x / (W - 1);
	for (int y = 0; y < H; ++y) (*htab)[(size_t)y] = (double)y / (H - 1);
}

// Spherical direction factors: dir = (sin_va*cos_ha, sin_va*sin_ha, cos_va) (src/Spherical.cpp:22-26).
// Layout of `out`: cos_ha[W], sin_ha[W], sin_va[H], cos_va[H].
inline void fill_spherical_tables(const PlaneConst &pc, int W, int H, const std::vector<double> &wtab,
                                  const std::vector<double> &htab, std::vector<double> *out) {
	out->resize((size_t)(2 * W + 2 * H));
	double *cos_ha = out->data(), *sin_ha = cos_ha + W, *sin_va = sin_ha + W, *cos_va = sin_va + H;
	for (int x = 0; x < W; ++x) {
		const double ha = pc.ul_hang - wtab[(size_t)x] * pc.hfov;
		cos_ha[x] = std::cos(ha);
		sin_ha[x] = std::sin(ha);
	}
	for (int y = 0; y < H; ++y) {
		const double va = pc.ul_vang + htab[(size_t)y] * pc.vfov;
		sin_va[y] = std::sin(va);
		cos_va[y] = std::cos(va);
	}
}

} // namespace hmrm

#endif
