/* r3d_oracle.c -- TEST INFRASTRUCTURE ONLY (never part of the product path).
 *
 * Plain-C, CPU restatement of the Radiative3D per-phonon propagate loop, on the
 * flattened r3d_model_desc and the Philox4x32-10 draw stream the CUDA path uses.
 * Every function cites the reference file:line it restates.  Arithmetic is kept
 * in the reference's operation order (and the file is compiled with
 * -ffp-contract=off) so that, on the same draws, this code and the reference's
 * own Propagate() agree to the last few ulps and can be compared phonon by
 * phonon (tests/test_oracle_vs_reference.py).
 *
 * PARITY PIN: golden vectors in tests/golden/ made by oracle/ref_harness.cpp from
 * the reference's compiled objects (incl. its --rtcoef-test table), plus, when
 * oracle/_ref is built, whole runs of the reference loop on the same Philox stream.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <complex.h>
#include <pthread.h>
#include "r3d_oracle.h"

#define R3D_PI     3.14159265358979323846   /* geom_base.hpp:32 */
#define R3D_PI45   (R3D_PI * 0.25)
#define R3D_PI90   (R3D_PI * 0.5)
#define R3D_PI180  (R3D_PI)
#define R3D_PI270  (R3D_PI * 1.5)
#define R3D_PI360  (R3D_PI * 2.0)
#define R3D_RAND_MAX 2147483647.0
#define PINF (1.0 / 0.0)
#define NINF (-1.0 / 0.0)

typedef struct { double x, y, z; } v3;

/* ---- R3::XYZ (geom_r3.hpp:113-240) ------------------------------------- */
static inline v3 V(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static inline double dot(v3 a, v3 b) { return b.x * a.x + b.y * a.y + b.z * a.z; }      /* :206 */
static inline v3 cross(v3 a, v3 b) {                                                    /* :222 */
  return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 vto(v3 a, v3 b) { return V(b.x - a.x, b.y - a.y, b.z - a.z); }         /* VectorTo :210 */
static inline v3 scal(v3 a, double s) { return V(s * a.x, s * a.y, s * a.z); }          /* ScaledBy :230 */
static inline v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline double mag2(v3 a) { return a.x * a.x + a.y * a.y + a.z * a.z; }           /* :129 */
static inline double mag(v3 a) { return sqrt(mag2(a)); }
static inline int iszero(v3 a) { return a.x == 0.0 && a.y == 0.0 && a.z == 0.0; }       /* :137 */
static inline v3 normalize(v3 a) { double n = 1.0 / mag(a); return V(a.x * n, a.y * n, a.z * n); } /* :185 */
static inline v3 unit_else(v3 a, v3 fb) {                                               /* :155 */
  double m = mag(a);
  if (m == 0.0) return fb;
  double mi = 1.0 / m;
  return V(a.x * mi, a.y * mi, a.z * mi);
}
static inline double xyz_theta(v3 a) { return (mag2(a) == 0.0) ? 0.0 : acos(a.z / mag(a)); }  /* :119 */
static inline double xyz_phi(v3 a) { return atan2(a.y, a.x); }                          /* :125 */
/* R3::XYZ(const S2::ThetaPhi&)  geom_r3.cpp:41-45 */
static inline v3 from_thph(double th, double ph) {
  return V(sin(th) * cos(ph), sin(th) * sin(ph), cos(th));
}
/* R3::XYZ::ThetaHat / PhiHat  geom_r3.cpp:85-126 */
static v3 xyz_thetahat(v3 a) {
  double theta = xyz_theta(a), phi = xyz_phi(a), rth, rph;
  if (theta < R3D_PI90) { rth = R3D_PI90 + theta; rph = phi; }
  else { rth = R3D_PI270 - theta; rph = (phi < R3D_PI180) ? phi + R3D_PI180 : phi - R3D_PI180; }
  return V(sin(rth) * cos(rph), sin(rth) * sin(rph), cos(rth));
}
static v3 xyz_phihat(v3 a) {
  double rph = xyz_phi(a) + R3D_PI90;
  return V(cos(rph), sin(rph), 0);
}
/* S2::ThetaPhi::ThetaHat / PhiHat (geom_s2.cpp:165-186, geom_s2.hpp:235-242),
 * converted to XYZ as the implicit conversion in Dot() does */
static v3 thph_thetahat(double th, double ph) {
  double rth, rph;
  if (th < R3D_PI90) { rth = R3D_PI90 + th; rph = ph; }
  else { rth = R3D_PI270 - th; rph = (ph < R3D_PI180) ? ph + R3D_PI180 : ph - R3D_PI180; }
  return from_thph(rth, rph);
}
static v3 thph_phihat(double th, double ph) {
  (void)th;
  return from_thph(R3D_PI90, (ph < R3D_PI270) ? ph + R3D_PI90 : ph - R3D_PI270);
}
/* geom_r3.cpp:146-171 */
static v3 inplane_unit_perp(v3 self, v3 other) {
  v3 mp = cross(self, other);
  if (iszero(mp)) {
    mp = cross(self, V(1, 0, 0));
    if (iszero(mp)) mp = cross(self, V(0, 1, 0));
  }
  mp = normalize(mp);
  v3 r = cross(mp, self);
  return normalize(r);
}
/* S2::ThetaPhi(const Node&) after Node(x,y,z) normalisation
 * (geom_s2.hpp:130-133,202-205; geom_s2.cpp:340-351) */
static void thph_from_node(v3 a, double *th, double *ph) {
  if (!(a.x == 0 && a.y == 0 && a.z == 0)) {
    double n = sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    a.x /= n; a.y /= n; a.z /= n;
  }
  *th = acos(a.z);
  *ph = atan2(a.y, a.x);
}

/* ---- R3::OrthoAxes (geom_r3.cpp:212-233) -------------------------------- */
typedef struct { v3 s1, s2, e3; } axes;
static axes make_axes(double the, double phi, double rot) {
  double ct = cos(the), st = sin(the), cp = cos(phi), sp = sin(phi), cr = cos(rot), sr = sin(rot);
  axes A;
  A.e3 = V(st * cp, st * sp, ct);
  A.s1 = V(cr * ct * cp - sr * sp, cr * ct * sp + sr * cp, -cr * st);
  A.s2 = V(-sr * ct * cp - cr * sp, -sr * ct * sp + cr * cp, sr * st);
  return A;
}
/* OrthoAxes::Express(XYZ)  geom_r3.hpp:560-568 */
static inline v3 axes_express(const axes *A, v3 v) {
  return V(v.x * A->s1.x + v.y * A->s2.x + v.z * A->e3.x,
           v.x * A->s1.y + v.y * A->s2.y + v.z * A->e3.y,
           v.x * A->s1.z + v.y * A->s2.z + v.z * A->e3.z);
}

/* Phonon::Transform (phonons.cpp:116-170) with OrthoAxes::Express(OrthoAxes)
 * (geom_r3.cpp:241-300) */
static void transform(double *th, double *ph, double *pol, double rth, double rph, double rpol) {
  axes AA = make_axes(*th, *ph, *pol);
  axes BB = make_axes(rth, rph, rpol);
  v3 s1 = axes_express(&AA, BB.s1);
  v3 e3 = axes_express(&AA, BB.e3);
  double costhe = e3.z;
  double the = acos(costhe);
  double phi = atan2(e3.y, e3.x);
  double sinthe = sin(the), cosphi = cos(phi), sinphi = sin(phi);
  v3 e1 = V(costhe * cosphi, costhe * sinphi, -sinthe);
  v3 e2 = V(-sinphi, cosphi, 0);
  double rot_x = dot(s1, e1), rot_y = dot(s1, e2);
  *th = the; *ph = phi; *pol = atan2(rot_y, rot_x);
}

/* Phonon::DirectionOfMotion (phonons.cpp:201-211) */
static v3 dir_of_motion(int type, double th, double ph, double pol) {
  if (type == R3D_RAY_P) return from_thph(th, ph);
  axes A = make_axes(th, ph, pol);
  return A.s1;
}

/* ---- Philox4x32-10: counter (idx_lo, idx_hi, block, 0), key (seed_lo, seed_hi) */
static void philox4x32_10(uint64_t seed, uint64_t idx, uint32_t block, uint32_t out[4]) {
  uint32_t c0 = (uint32_t)idx, c1 = (uint32_t)(idx >> 32), c2 = block, c3 = 0u;
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int r = 0; r < 10; r++) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
uint32_t r3d_oracle_draw(uint64_t seed, uint64_t idx, uint32_t ordinal) {
  uint32_t w[4];
  philox4x32_10(seed, idx, ordinal >> 2, w);
  return w[ordinal & 3] >> 1;                    /* 31 bits == [0, RAND_MAX] */
}
typedef struct { uint64_t seed, idx; uint32_t ordinal; } rng_t;
static inline uint32_t rng_next(rng_t *g) { return r3d_oracle_draw(g->seed, g->idx, g->ordinal++); }

/* ProbDist::GetRandomIndex (probability.cpp:104-129) */
static uint32_t cdf_search(const double *cdf, uint32_t n, uint32_t kdraw) {
  uint32_t k1 = 0, k2 = n - 1, k;
  double r = cdf[k2] * ((double)kdraw / R3D_RAND_MAX);
  while (k1 != k2) {
    k = (k1 + k2) >> 1;
    if (r <= cdf[k]) k2 = k; else k1 = k + 1;
  }
  return k2;
}

/* ---- travel record (media.hpp:94-108) ----------------------------------- */
typedef struct { double len, time; v3 loc; double th, ph, atten; int face; } travel_t;

/* MediumCell::HelperUniformAttenuation (media.cpp:98-100) */
static inline double atten_uniform(double cycles, double Q) { return exp(((-1) * R3D_PI * cycles) / Q); }

/* PlaneFace::LinearRayDistToExit (media_cellface.cpp:262-324) */
static double plane_dist_exit(v3 N, v3 P, v3 loc, v3 dir) {
  double d_sh = dot(N, vto(loc, P));
  double d_fact = dot(N, dir);
  if (d_fact < 0) return PINF;
  if (d_fact == 0) return (d_sh < 0) ? NINF : PINF;
  return d_sh / d_fact;
}
/* CylinderFace::LinearRayDistToExit (media_cellface.cpp:531-562) */
static double cyl_dist_exit(double rad2, v3 loc, v3 dir) {
  double A = dir.x * dir.x + dir.y * dir.y;
  double C = loc.x * loc.x + loc.y * loc.y - rad2;
  if (A == 0) return (C <= 0) ? PINF : NINF;
  double B = 2 * (loc.x * dir.x + loc.y * dir.y);
  double urad = B * B - 4 * A * C;
  if (urad < 0) return NINF;
  double margin = sqrt(urad);
  return (margin - B) / (2 * A);
}
/* SphereFace::LinearRayDistToExit (media_cellface.cpp:664-684); outward == (signed radius > 0) */
static double sphere_dist_exit(double rad2, int outward, v3 loc, v3 dir) {
  double midpt = -dot(loc, dir);
  double urad = rad2 + midpt * midpt - mag2(loc);
  if (urad <= 0) return outward ? NINF : PINF;
  double sqrad = sqrt(urad);
  double dplus = midpt + sqrad;
  if (outward) return dplus;
  if (midpt <= 0) return PINF;
  return midpt - sqrad;
}

/* ---- RCUCylinder (media.cpp:208-330) ------------------------------------ */
static travel_t cyl_advance(const r3d_model_desc *d, const double *c, int rt, double len, v3 loc, double th, double ph) {
  travel_t r;
  r.len = len;
  r.time = len / c[rt];
  r.loc = add(loc, scal(from_thph(th, ph), len));
  r.th = th; r.ph = ph;
  r.atten = atten_uniform(r.time * d->freq_hz, c[3 + rt]);
  r.face = -1;
  return r;
}
static travel_t cyl_path(const r3d_model_desc *d, const double *c, int rt, v3 loc, double th, double ph) {
  v3 dir = from_thph(th, ph);
  double dl = cyl_dist_exit(d->cyl_radius2, loc, dir);
  double dt = plane_dist_exit(V(c[5], c[6], c[7]), V(c[8], c[9], c[10]), loc, dir);
  double db = plane_dist_exit(V(c[11], c[12], c[13]), V(c[14], c[15], c[16]), loc, dir);
  if (dl < 0) dl = 0;
  if (dt < 0) dt = 0;
  if (db < 0) db = 0;
  int exf = 2; double shortest = dl;
  if (dt < shortest) { exf = 0; shortest = dt; }
  if (db < shortest) { exf = 1; shortest = db; }
  travel_t r = cyl_advance(d, c, rt, shortest, loc, th, ph);
  r.face = exf;
  return r;
}

/* ---- SphereShell (media.cpp:668-970), RayArcAttributes (raypath.hpp:31-113, raypath.cpp:5-19) */
typedef struct {
  double radius, rad2; v3 center, u3, u2, u1;
  double S, S2, TwoSQ, CosZeta, SinZeta, CotZetaBy2, timeCoef;
} arc_t;
/* EarthCoords::GetDown (ecs.hpp:373, ecs.cpp:147-167); singular up = (0,1,0) (ecs.cpp:119) */
static v3 ecs_down(const r3d_model_desc *d, v3 loc) {
  if (!d->ecs_radial) return neg(V(0, 0, +1));
  v3 c = V(d->earth_center[0], d->earth_center[1], d->earth_center[2]);
  return neg(unit_else(vto(c, loc), V(0, 1, 0)));
}
static inline double shell_veloc(const double *c, int rt, v3 loc) { return c[2 + rt] + c[rt] * mag2(loc); } /* :646 */
static arc_t shell_ray_arc(const r3d_model_desc *d, const double *c, int rt, v3 loc, double th, double ph) {  /* :779-840 */
  arc_t R;
  v3 dir = from_thph(th, ph);
  v3 v3_ = ecs_down(d, loc);
  v3 v2 = unit_else(cross(v3_, dir), V(0, 0, 0));
  v3 v1 = cross(v2, v3_);
  double sini = dot(v1, dir);
  if (sini > 1.0) sini = 1.0;
  double cosi = dot(v3_, dir);
  const double G = sini * mag(loc) / shell_veloc(c, rt, loc);
  const double TwoGA = 2. * G * c[rt];
  const double urad = 1. - (2. * TwoGA * G * c[2 + rt]);
  double Bottom = (urad > 1) ? (1. - sqrt(urad)) / TwoGA : 0;
  R.radius = (c[4 + rt] / Bottom - Bottom) / 2.0;
  R.rad2 = R.radius * R.radius;
  v3 l2c = add(scal(v1, R.radius * cosi), scal(v3_, -R.radius * sini));
  R.center = add(loc, l2c);
  R.u3 = ecs_down(d, R.center);
  R.u2 = v2;
  R.u1 = cross(R.u2, R.u3);
  if (urad <= 1) {
    R.center = V(0, 0, 0);
    R.u3 = R.u2 = V(0, 0, 0);
    R.u1 = dir;
  }
  /* cache_RD2_precompute (raypath.hpp:42-52) */
  R.S2 = mag2(R.center);
  R.S = sqrt(R.S2);
  R.TwoSQ = 2 * R.S * R.radius;
  R.CosZeta = (R.S2 + R.radius * R.radius - c[4 + rt]) / R.TwoSQ;
  R.SinZeta = sqrt(1 - R.CosZeta * R.CosZeta);
  R.CotZetaBy2 = (1 + R.CosZeta) / R.SinZeta;
  R.timeCoef = -1 / (c[rt] * R.S * R.SinZeta);
  return R;
}
static inline double arc_angle_from_bottom(const arc_t *a, v3 loc) {   /* raypath.cpp:5-10 */
  v3 c2l = vto(a->center, loc);
  return atan2(dot(a->u1, c2l), dot(a->u3, c2l));
}
/* SphereFace::CircularArcDistToExit (media_cellface.cpp:717-748) */
static double sphere_arc_dist_exit(double rad2, int outward, v3 loc, v3 dir, const arc_t *arc) {
  if (arc->S2 == 0) return sphere_dist_exit(rad2, outward, loc, dir);
  double cosq = (arc->S2 + arc->rad2 - rad2) / arc->TwoSQ;
  if (cosq > 1.0) return outward ? NINF : PINF;
  double angleBtoE = acos(cosq);
  double angleLoc = arc_angle_from_bottom(arc, loc);
  double angleLtoE = angleBtoE - angleLoc;
  if (outward) return angleLtoE * arc->radius;
  if (angleLoc >= 0) return PINF;
  angleLtoE = -angleBtoE - angleLoc;
  return angleLtoE * arc->radius;
}
static travel_t shell_advance_rd0(const r3d_model_desc *d, const double *c, int rt, double len, v3 loc, double th, double ph) { /* :869-880 */
  travel_t r;
  r.len = len;
  r.time = len / c[2 + rt];
  r.loc = add(loc, scal(from_thph(th, ph), len));
  r.th = th; r.ph = ph;
  r.atten = atten_uniform(r.time * d->freq_hz, c[8 + rt]);
  r.face = -1;
  return r;
}
static travel_t shell_advance_rd2_impl(const r3d_model_desc *d, const double *c, int rt, double len, v3 loc,
                                       double th, double ph, const arc_t *arc) {   /* :905-960 */
  if (arc->radius == PINF) {
    travel_t fb = shell_advance_rd0(d, c, rt, len, loc, th, ph);
    double r0 = mag(loc), r1 = mag(fb.loc);
    double sqnac = sqrt(-c[rt] * c[2 + rt]);
    double sqnaoc = sqrt(-c[rt] / c[2 + rt]);
    double tau0 = atanh(sqnaoc * r0), tau1 = atanh(sqnaoc * r1);
    double tpm = (tau1 - tau0) / sqnac;
    fb.time = fabs(tpm);           /* media.cpp:927 unqualified abs -> double overload */
    return fb;
  }
  double startAngle = arc_angle_from_bottom(arc, loc);
  double angleDelta = len / arc->radius;
  double endAngle = startAngle + angleDelta;
  /* PositionFromAngle / DirectionFromAngle (raypath.cpp:11-19) */
  v3 newLoc = add(add(arc->center, scal(arc->u1, arc->radius * sin(endAngle))), scal(arc->u3, arc->radius * cos(endAngle)));
  v3 newDir = add(scal(arc->u1, cos(endAngle)), scal(arc->u3, -sin(endAngle)));
  /* GetTravelTimeAngleToAngle_RD2 (media.cpp:962-970) */
  double t0 = arc->timeCoef * atanh(arc->CotZetaBy2 * tan(startAngle / 2));
  double t1 = arc->timeCoef * atanh(arc->CotZetaBy2 * tan(endAngle / 2));
  double timeDelta = t1 - t0;
  travel_t r;
  r.len = len; r.time = timeDelta; r.loc = newLoc;
  thph_from_node(newDir, &r.th, &r.ph);
  r.atten = atten_uniform(timeDelta * d->freq_hz, c[8 + rt]);
  r.face = -1;
  return r;
}
static travel_t shell_advance(const r3d_model_desc *d, const double *c, int rt, double len, v3 loc, double th, double ph) { /* :859-867,893-903 */
  if (c[rt] == 0) return shell_advance_rd0(d, c, rt, len, loc, th, ph);
  arc_t arc = shell_ray_arc(d, c, rt, loc, th, ph);
  return shell_advance_rd2_impl(d, c, rt, len, loc, th, ph, &arc);
}
static travel_t shell_path(const r3d_model_desc *d, const double *c, int rt, v3 loc, double th, double ph) { /* :668-760 */
  v3 dir = from_thph(th, ph);
  int out0 = c[10] > 0, out1 = c[11] > 0;
  double dists[2];
  if (c[rt] < 0) {
    arc_t arc = shell_ray_arc(d, c, rt, loc, th, ph);
    dists[0] = sphere_arc_dist_exit(c[12], out0, loc, dir, &arc);
    dists[1] = sphere_arc_dist_exit(c[13], out1, loc, dir, &arc);
    int ef = (dists[0] < dists[1]) ? 0 : 1;
    if (dists[ef] < 0) dists[ef] = 0;
    travel_t r = shell_advance_rd2_impl(d, c, rt, dists[ef], loc, th, ph, &arc);
    r.face = ef;
    return r;
  }
  /* a == 0: straight rays (a > 0 is rejected by the reference with an exception, media.cpp:675) */
  dists[0] = sphere_dist_exit(c[12], out0, loc, dir);
  dists[1] = sphere_dist_exit(c[13], out1, loc, dir);
  int ef = (dists[0] < dists[1]) ? 0 : 1;
  if (dists[ef] < 0) dists[ef] = 0;
  travel_t r = shell_advance_rd0(d, c, rt, dists[ef], loc, th, ph);
  r.face = ef;
  return r;
}

/* ---- Tetra (media.cpp:412-567), CoordinateTransformation (media.hpp:549-598) */
typedef struct { v3 r1, r2, r3; } m3;   /* rows */
static inline v3 m3_mul(const m3 *M, v3 v) {            /* geom_r3.hpp:365-369 */
  return V((M->r1.x * v.x) + (M->r1.y * v.y) + (M->r1.z * v.z),
           (M->r2.x * v.x) + (M->r2.y * v.y) + (M->r2.z * v.z),
           (M->r3.x * v.x) + (M->r3.y * v.y) + (M->r3.z * v.z));
}
static inline v3 m3_tmul(const m3 *M, v3 v) {           /* M.T() * v */
  return V((M->r1.x * v.x) + (M->r2.x * v.y) + (M->r3.x * v.z),
           (M->r1.y * v.x) + (M->r2.y * v.y) + (M->r3.y * v.z),
           (M->r1.z * v.x) + (M->r2.z * v.y) + (M->r3.z * v.z));
}
typedef struct { v3 prime, trans; m3 S; double R; } ct_t;
static ct_t tetra_ct(double Vxo, v3 g, v3 loc, v3 t) {
  ct_t C;
  v3 v2 = cross(g, t);
  v3 v1 = cross(v2, g);
  v3 v3_ = g;
  v1 = normalize(v1); v2 = normalize(v2); v3_ = normalize(v3_);
  double txprime = dot(t, v1);
  double tzprime = dot(t, v3_);
  double s = txprime / (Vxo);
  double R = 1 / (s * mag(g));
  C.S.r1 = v1; C.S.r2 = v2; C.S.r3 = v3_;
  v3 x0rot = m3_mul(&C.S, loc);
  v3 translate = V(x0rot.x + R * tzprime, x0rot.y, x0rot.z + (-1) * R * txprime);
  C.prime = add(x0rot, scal(translate, -1));
  C.trans = translate;
  C.R = R;
  return C;
}
static inline double tetra_veloc(const double *c, int rt, v3 loc) {      /* :412-414 */
  return dot(loc, V(c[3 * rt], c[3 * rt + 1], c[3 * rt + 2])) + c[6 + rt];
}
static travel_t tetra_advance(const r3d_model_desc *d, const double *c, int rt, double len, v3 loc, double th, double ph) { /* :442-499 */
  v3 g = V(c[3 * rt], c[3 * rt + 1], c[3 * rt + 2]);
  ct_t CT = tetra_ct(tetra_veloc(c, rt, loc), g, loc, from_thph(th, ph));
  double theta = len / CT.R;
  v3 nlr = V(CT.R * sin(theta / 2), 0, CT.R * cos(theta / 2));
  double angletoX0 = atan2(CT.prime.x, CT.prime.z);
  double rotAngle = angletoX0 + (theta / 2);
  rotAngle = (rotAngle > R3D_PI360) ? rotAngle - R3D_PI360 : rotAngle;
  v3 nl2 = V(cos(rotAngle) * nlr.x + sin(rotAngle) * nlr.z, 0, -sin(rotAngle) * nlr.x + cos(rotAngle) * nlr.z);
  v3 newLoc = m3_tmul(&CT.S, add(nl2, CT.trans));
  double AngleNewLoc2D = atan2(nl2.x, nl2.z);
  v3 nd2 = V(cos(AngleNewLoc2D), 0, (-1) * sin(AngleNewLoc2D));
  v3 newDir = normalize(m3_tmul(&CT.S, nd2));
  double travelTime = (1 / mag(g)) * (log(fabs(tan((AngleNewLoc2D / 2 + R3D_PI45))))
                                      - log(fabs(tan((angletoX0 / 2 + R3D_PI45)))));
  travel_t r;
  r.len = len; r.time = travelTime; r.loc = newLoc;
  thph_from_node(newDir, &r.th, &r.ph);
  r.atten = atten_uniform(travelTime * d->freq_hz, c[12 + rt]);
  r.face = -1;
  return r;
}
typedef struct { double entry, exit, half; int cont; } gcad_t;
/* PlaneFace::GetCircArcDistToFace (media_cellface.cpp:333-426) */
static gcad_t tetra_gcad(v3 N, v3 P, double R, v3 C, const m3 *S) {
  gcad_t g;
  int continuous = 1;
  v3 rotNorm = m3_mul(S, N);
  v3 rotx0 = m3_mul(S, P);
  v3 x0prime = add(rotx0, scal(C, -1));
  double d = (-1) * dot(rotNorm, x0prime);
  double D = -d / sqrt(rotNorm.x * rotNorm.x + rotNorm.z * rotNorm.z);
  v3 norm2D = normalize(V(rotNorm.x, 0, rotNorm.z));
  double bis = atan2(norm2D.x, norm2D.z);
  double ex = 0, en = 0;
  if (D / R < 1 && D / R > -1) {
    double a = acos(D / R);
    if (bis > (-1) * R3D_PI90 && bis < R3D_PI90) { en = bis + a; ex = bis - a; continuous = 0; }
    else if (bis <= (-1) * R3D_PI90) { en = bis + a; ex = bis - a + R3D_PI360; }
    else if (bis >= R3D_PI90) { en = bis + a - R3D_PI360; ex = bis - a; }
    else { en = ex = bis = NAN; }          /* reference exit(1)s here ("GCAD Bisector nan") */
  }
  if (bis >= R3D_PI90 || bis <= (-1) * R3D_PI90) bis = PINF;
  if (en >= R3D_PI90) en = PINF;
  if (en <= (-1) * R3D_PI90) en = NINF;
  if (ex >= R3D_PI90) ex = PINF;
  if (ex <= (-1) * R3D_PI90) ex = NINF;
  if (D / R >= 1) { en = NINF; ex = PINF; }
  if (D / R <= -1) { en = PINF; ex = NINF; bis = NINF; continuous = 0; }
  g.entry = en; g.exit = ex; g.half = bis; g.cont = continuous;
  return g;
}
/* GCAD_RetVal::Inside / IsProper (media_cellface.cpp:767-794) */
static int gcad_inside(const gcad_t *g, double theta) {
  double error = 0.0000000001;
  if (g->cont) { if (theta <= g->exit && theta >= (g->entry - error)) return 1; }
  if (!g->cont) {
    if ((theta >= (-1) * R3D_PI90 && theta <= g->exit) || (theta >= (g->entry - error) && theta <= R3D_PI90)) return 1;
  }
  return 0;
}
static travel_t tetra_path(const r3d_model_desc *d, const double *c, int rt, v3 loc, double th, double ph) { /* :518-567 */
  v3 g = V(c[3 * rt], c[3 * rt + 1], c[3 * rt + 2]);
  ct_t CT = tetra_ct(tetra_veloc(c, rt, loc), g, loc, from_thph(th, ph));
  double ColatAngletoX0 = atan2(CT.prime.x, CT.prime.z);
  gcad_t rv[4];
  for (int i = 0; i < 4; i++) {
    const double *f = c + 14 + 6 * i;
    rv[i] = tetra_gcad(V(f[0], f[1], f[2]), V(f[3], f[4], f[5]), CT.R, CT.trans, &CT.S);
  }
  double len = PINF;
  int faceID = 0;
  for (int i = 0; i < 4; i++) {
    double ex = rv[i].exit;
    if (gcad_inside(&rv[(i + 1) % 4], ex) && gcad_inside(&rv[(i + 2) % 4], ex) && gcad_inside(&rv[(i + 3) % 4], ex)) {
      double newlen = (ex - ColatAngletoX0) * CT.R;
      if (newlen < 0 && (ColatAngletoX0 > rv[i].half)) newlen = len;
      if (newlen < len) { len = newlen; faceID = i; }
    }
  }
  travel_t r = tetra_advance(d, c, rt, len, loc, th, ph);
  r.face = faceID;
  return r;
}

/* ---- dispatch on cell kind ---------------------------------------------- */
static inline const double *cellp(const r3d_model_desc *d, uint32_t cell) { return d->cell_params + (size_t)cell * d->cell_nparam; }
static travel_t path_to_boundary(const r3d_model_desc *d, uint32_t cell, int rt, v3 loc, double th, double ph) {
  const double *c = cellp(d, cell);
  switch (d->cell_kind) {
  case R3D_CELL_CYLINDER: return cyl_path(d, c, rt, loc, th, ph);
  case R3D_CELL_SHELL:    return shell_path(d, c, rt, loc, th, ph);
  default:                return tetra_path(d, c, rt, loc, th, ph);
  }
}
static travel_t advance_length(const r3d_model_desc *d, uint32_t cell, int rt, double len, v3 loc, double th, double ph) {
  const double *c = cellp(d, cell);
  switch (d->cell_kind) {
  case R3D_CELL_CYLINDER: return cyl_advance(d, c, rt, len, loc, th, ph);
  case R3D_CELL_SHELL:    return shell_advance(d, c, rt, len, loc, th, ph);
  default:                return tetra_advance(d, c, rt, len, loc, th, ph);
  }
}
/* GetVelocAtPoint / GetDensityAtPoint (media.cpp:185-196, 412-425, 646-660) */
static double veloc_at(const r3d_model_desc *d, uint32_t cell, int rt, v3 loc) {
  const double *c = cellp(d, cell);
  switch (d->cell_kind) {
  case R3D_CELL_CYLINDER: return c[rt];
  case R3D_CELL_SHELL:    return shell_veloc(c, rt, loc);
  default:                return tetra_veloc(c, rt, loc);
  }
}
static double dens_at(const r3d_model_desc *d, uint32_t cell, v3 loc) {
  const double *c = cellp(d, cell);
  switch (d->cell_kind) {
  case R3D_CELL_CYLINDER: return c[2];
  case R3D_CELL_SHELL:    return c[7] + c[6] * mag2(loc);
  default:                return dot(loc, V(c[8], c[9], c[10])) + c[11];
  }
}
/* CellFace::Normal (media_cellface.hpp:298; media_cellface.cpp:506-510, 594-597) */
static v3 face_normal(const r3d_model_desc *d, uint32_t cell, int face, v3 loc) {
  const double *c = cellp(d, cell);
  switch (d->cell_kind) {
  case R3D_CELL_CYLINDER:
    if (face == 0) return V(c[5], c[6], c[7]);
    if (face == 1) return V(c[11], c[12], c[13]);
    return unit_else(V(loc.x, loc.y, 0), V(1, 0, 0));
  case R3D_CELL_SHELL: {
    v3 u = unit_else(loc, V(0, 0, 1));
    return (c[10 + face] > 0) ? u : neg(u);
  }
  default: { const double *f = c + 14 + 6 * face; return V(f[0], f[1], f[2]); }
  }
}

/* ---- RTCoef (rtcoef.cpp:30-588) ------------------------------------------ */
enum { R_P = 0, R_SV, R_SH, T_P, T_SV, T_SH, RT_NUM };
typedef double complex cx;
typedef struct {
  int notransmit; v3 fnorm, fpara, fparash; double sini;
  double densR, densT, velR[2], velT[2];
  double sino[RT_NUM]; cx coso[RT_NUM]; cx amp[RT_NUM]; double prob[RT_NUM];
  int defchoice, choice; v3 chosen_dir;
} rtcoef_t;
static inline cx csqrt_real(double x) { return csqrt(x + 0.0 * I); }
static inline double cnorm(cx a) { return creal(a) * creal(a) + cimag(a) * cimag(a); }   /* std::norm */
static void rt_init(rtcoef_t *rt, v3 fnorm, v3 phdir) {                /* rtcoef.cpp:30-52 */
  memset(rt, 0, sizeof *rt);
  rt->fnorm = fnorm;
  rt->fpara = inplane_unit_perp(fnorm, phdir);
  rt->fparash = cross(fnorm, rt->fpara);
  rt->sini = dot(rt->fpara, phdir);
}
static void rt_coefs_psv(rtcoef_t *rt, int intype) {                   /* rtcoef.cpp:107-205, 289-404 */
  const double rho1 = rt->densR, rho2 = rt->densT;
  const double alpha1 = rt->velR[0], alpha2 = rt->velT[0], beta1 = rt->velR[1], beta2 = rt->velT[1];
  const int rtyp = (intype == R3D_RAY_P) ? 0 : 1;
  const double p = rt->sini / rt->velR[rtyp];
  rt->sino[T_P] = rt->velT[0] * p;
  rt->sino[T_SV] = rt->velT[1] * p;
  rt->sino[R_SV] = rt->velR[1] * p;
  rt->sino[R_P] = rt->velR[0] * p;
  rt->coso[T_P] = csqrt_real(1.0 - rt->sino[T_P] * rt->sino[T_P]);
  rt->coso[T_SV] = csqrt_real(1.0 - rt->sino[T_SV] * rt->sino[T_SV]);
  rt->coso[R_SV] = csqrt_real(1.0 - rt->sino[R_SV] * rt->sino[R_SV]);
  rt->coso[R_P] = csqrt_real(1.0 - rt->sino[R_P] * rt->sino[R_P]);
  double beta1_sq = beta1 * beta1, beta2_sq = beta2 * beta2, p_sq = p * p;
  double tmp1 = rho1 * (1. - 2. * beta1_sq * p_sq);
  double tmp2 = rho2 * (1. - 2. * beta2_sq * p_sq);
  double tmp3 = 2. * rho1 * beta1_sq;
  double tmp4 = 2. * rho2 * beta2_sq;
  const double a = tmp2 - tmp1, b = tmp2 + tmp3 * p_sq, c = tmp1 + tmp4 * p_sq, dd = tmp4 - tmp3;
  const cx cosi1 = rt->coso[R_P] / alpha1, cosi2 = rt->coso[T_P] / alpha2;
  const cx cosj1 = rt->coso[R_SV] / beta1, cosj2 = rt->coso[T_SV] / beta2;
  const cx E = b * cosi1 + c * cosi2;
  const cx F = b * cosj1 + c * cosj2;
  const cx G = a - dd * cosi1 * cosj2;
  const cx H = a - dd * cosi2 * cosj1;
  const cx D = E * F + G * H * p_sq;
  const double two = 2.0;
  cx Term1, Term2;
  if (intype == R3D_RAY_P) {
    Term1 = ((b * cosi1) - (c * cosi2));
    Term2 = ((a) + (dd * cosi1 * cosj2));
    rt->amp[R_P] = (Term1 * F - Term2 * H * p_sq) / D;
    Term1 = (a * b + c * dd * cosi2 * cosj2);
    rt->amp[R_SV] = -two * cosi1 * Term1 * p * alpha1 / (beta1 * D);
    Term1 = two * rho1 * cosi1 * alpha1;
    rt->amp[T_P] = Term1 * F / (alpha2 * D);
    rt->amp[T_SV] = Term1 * H * p / (beta2 * D);
  } else {
    Term1 = (a * b + c * dd * cosi2 * cosj2);
    rt->amp[R_P] = -two * cosj1 * Term1 * p * beta1 / (alpha1 * D);
    Term1 = (b * cosj1 - c * cosj2);
    Term2 = (a + dd * cosi2 * cosj1);
    rt->amp[R_SV] = -(Term1 * E - Term2 * G * p_sq) / D;
    Term1 = two * rho1 * cosj1 * beta1;
    rt->amp[T_P] = -Term1 * G * p / (alpha2 * D);
    rt->amp[T_SV] = Term1 * E / (beta2 * D);
  }
  rt->prob[R_SH] = 0; rt->prob[T_SH] = 0;
  rt->prob[R_P] = rho1 * alpha1 * creal(rt->coso[R_P]) * cnorm(rt->amp[R_P]);
  rt->prob[R_SV] = rho1 * beta1 * creal(rt->coso[R_SV]) * cnorm(rt->amp[R_SV]);
  rt->prob[T_P] = rho2 * alpha2 * creal(rt->coso[T_P]) * cnorm(rt->amp[T_P]);
  rt->prob[T_SV] = rho2 * beta2 * creal(rt->coso[T_SV]) * cnorm(rt->amp[T_SV]);
}
static void rt_coefs_sh(rtcoef_t *rt) {                                /* rtcoef.cpp:207-287 */
  rt->prob[R_P] = rt->prob[R_SV] = rt->prob[T_P] = rt->prob[T_SV] = 0;
  const double rho1 = rt->densR, rho2 = rt->densT, beta1 = rt->velR[1], beta2 = rt->velT[1];
  rt->sino[R_SH] = rt->sini;
  rt->sino[T_SH] = (beta2 / beta1) * rt->sini;
  rt->coso[R_SH] = csqrt_real(1.0 - rt->sino[R_SH] * rt->sino[R_SH]);
  rt->coso[T_SH] = csqrt_real(1.0 - rt->sino[T_SH] * rt->sino[T_SH]);
  cx a = rho1 * beta1 * rt->coso[R_SH];
  cx b = rho2 * beta2 * rt->coso[T_SH];
  rt->amp[R_SH] = (a - b) / (a + b);
  rt->amp[T_SH] = 2.0 * a / (a + b);
  rt->prob[R_SH] = rho1 * beta1 * creal(rt->coso[R_SH]) * cnorm(rt->amp[R_SH]);
  rt->prob[T_SH] = rho2 * beta2 * creal(rt->coso[T_SH]) * cnorm(rt->amp[T_SH]);
}
static void rt_get_coefs(rtcoef_t *rt, int intype) {                   /* rtcoef.cpp:76-105 */
  if (intype == R3D_RAY_P) { rt->defchoice = R_P; rt_coefs_psv(rt, R3D_RAY_P); }
  else if (intype == R3D_RAY_SH) { rt->defchoice = R_SH; rt_coefs_sh(rt); }
  else { rt->defchoice = R_SV; rt_coefs_psv(rt, R3D_RAY_SV); }
}
static int rt_choose_spol(const rtcoef_t *rt, v3 pdom, uint32_t k) {   /* rtcoef.cpp:406-423 */
  double shfrac = dot(pdom, rt->fparash);
  shfrac *= shfrac;
  double ran = ((double)k / R3D_RAND_MAX);
  return (ran <= shfrac) ? R3D_RAY_SH : R3D_RAY_SV;
}
static void rt_choose(rtcoef_t *rt, uint32_t k) {                       /* rtcoef.cpp:436-475 */
  double PI[RT_NUM], TotalP;
  PI[0] = rt->prob[0];
  for (int i = 1; i < RT_NUM; i++) PI[i] = PI[i - 1] + rt->prob[i];
  TotalP = PI[RT_NUM - 1];
  if (k == 0) k = 1;
  double ran = ((double)k / R3D_RAND_MAX) * TotalP;
  int choice = RT_NUM - 1;
  for (int i = 0; i < RT_NUM - 1; i++) if (ran <= PI[i]) { choice = i; break; }
  if ((TotalP == 0) || ((TotalP - TotalP) != 0)) choice = rt->defchoice;
  if (rt->notransmit) {
    if (choice == T_P) choice = R_P;
    if (choice == T_SV) choice = R_SV;
    if (choice == T_SH) choice = R_SH;
  }
  rt->choice = choice;
}
static v3 rt_chosen_dir(rtcoef_t *rt) {                                 /* rtcoef.cpp:521-548 */
  double comp_para = rt->sino[rt->choice];
  double comp_norm = creal(rt->coso[rt->choice]);
  if (comp_para > 1.0) comp_para = 1.0;
  if (rt->choice == R_P || rt->choice == R_SV || rt->choice == R_SH) comp_norm *= -1;
  rt->chosen_dir = add(scal(rt->fpara, comp_para), scal(rt->fnorm, comp_norm));
  return rt->chosen_dir;
}
static v3 rt_chosen_pdom(const rtcoef_t *rt) {                          /* rtcoef.cpp:561-588 */
  if (rt->choice == T_P || rt->choice == R_P) return rt->chosen_dir;
  if (rt->choice == T_SH || rt->choice == R_SH) return rt->fparash;
  if (rt->choice == R_SV) return cross(rt->chosen_dir, rt->fparash);
  return cross(rt->fparash, rt->chosen_dir);
}

/* ---- Seismometer::CatchPhonon (dataout.cpp:103-216) ---------------------- */
/* returns 1 and fills bin/e[4] when the phonon is binned */
static int seis_catch(const double *s, double bin_dt, uint32_t n_bins, double time, v3 loc, double th, double ph,
                      double pol, int type, double amp, double vel, uint32_t *bin, double e[4]) {
  int within_window = 1, within_radius = 1;
  v3 sloc = V(s[0], s[1], s[2]);
  double arv = time, correction = 0;
  if (s[12 + type] <= 0) {
    v3 toseis = vto(loc, sloc);
    correction = dot(toseis, from_thph(th, ph));
    correction = correction / vel;
  }
  arv += correction;
  double scaled = ((arv - 0.0) / bin_dt);
  if (scaled < 0.0) within_window = 0;
  double fl = floor(scaled);
  if (!(fl < (double)n_bins)) within_window = 0;     /* unsigned binindex >= cmNumBins */
  double dist = mag(vto(loc, sloc));                 /* DistFrom, geom_r3.hpp:216 */
  if (dist > s[14 + type]) within_radius = 0;
  if (dist < s[12 + type]) within_radius = 0;
  if (!within_window) return 0;
  if (!within_radius) return 0;
  v3 dopm = dir_of_motion(type, th, ph, pol);
  double xf = dot(dopm, V(s[3], s[4], s[5])), yf = dot(dopm, V(s[6], s[7], s[8])), zf = dot(dopm, V(s[9], s[10], s[11]));
  xf *= xf; yf *= yf; zf *= zf;
  double energy = amp * amp;
  energy /= bin_dt;
  energy /= s[16 + type];
  e[0] = energy * xf; e[1] = energy * yf; e[2] = energy * zf; e[3] = energy;
  *bin = (uint32_t)fl;
  return 1;
}

/* ---- phonon + Propagate (phonons.cpp:540-682) ---------------------------- */
typedef struct {
  double time, pathlen, recent, amp; uint32_t moves;
  v3 loc; double th, ph, pol; int type; uint32_t cell;
  uint32_t catches, scatters, iters;
} phonon_t;

static inline void ph_move(phonon_t *p, const travel_t *t) {           /* phonons.cpp:62-70 */
  p->pathlen += t->len; p->time += t->time; p->recent += t->time;
  p->loc = t->loc; p->th = t->th; p->ph = t->ph;
  p->amp *= t->atten; p->moves += 1;
}
static inline double nudge(const r3d_model_desc *d, double th) {       /* phonons.hpp:335-344 */
  if (th < d->min_theta) th = d->min_theta;
  if (th > d->max_theta) th = d->max_theta;
  return th;
}

/* Phonon::Refraction_FullRT (phonons.cpp:429-476) + CellFace::GetRTBasis (media_cellface.cpp:122-149) */
static void refraction_fullrt(const r3d_model_desc *d, phonon_t *p, int face, rng_t *g) {
  uint32_t fi = p->cell * d->faces_per_cell + face;
  int adjoin = d->face_flags[fi] & R3D_FACE_ADJOIN;
  uint32_t other = d->face_other_cell[fi];
  rtcoef_t rt;
  rt_init(&rt, face_normal(d, p->cell, face, p->loc), from_thph(p->th, p->ph));
  rt.densR = dens_at(d, p->cell, p->loc);
  rt.velR[0] = veloc_at(d, p->cell, 0, p->loc);
  rt.velR[1] = veloc_at(d, p->cell, 1, p->loc);
  if (adjoin) {
    rt.densT = dens_at(d, other, p->loc);
    rt.velT[0] = veloc_at(d, other, 0, p->loc);
    rt.velT[1] = veloc_at(d, other, 1, p->loc);
  } else {
    rt.densT = 0.0; rt.velT[0] = 1e-12; rt.velT[1] = 1e-12; rt.notransmit = 1;
  }
  int intype = R3D_RAY_P;
  if (p->type == R3D_RAY_S) intype = rt_choose_spol(&rt, dir_of_motion(p->type, p->th, p->ph, p->pol), rng_next(g));
  rt_get_coefs(&rt, intype);
  rt_choose(&rt, rng_next(g));
  int outtype = (rt.choice == R_P || rt.choice == T_P) ? R3D_RAY_P : R3D_RAY_S;
  int transmit = !(rt.choice == R_P || rt.choice == R_SV || rt.choice == R_SH);
  v3 outdir = rt_chosen_dir(&rt);
  p->type = outtype;
  p->th = xyz_theta(outdir); p->ph = xyz_phi(outdir);
  if (p->type == R3D_RAY_S) {
    v3 pdomo = rt_chosen_pdom(&rt);
    double pol_v = dot(pdomo, thph_thetahat(p->th, p->ph));
    double pol_h = dot(pdomo, thph_phihat(p->th, p->ph));
    p->pol = atan2(pol_h, pol_v);
  }
  if (transmit) p->cell = other;
}
/* Phonon::Refraction_Bend (phonons.cpp:311-405) */
static void refraction_bend(const r3d_model_desc *d, phonon_t *p, int face) {
  uint32_t other = d->face_other_cell[p->cell * d->faces_per_cell + face];
  v3 mdir = from_thph(p->th, p->ph);
  v3 fnorm = face_normal(d, p->cell, face, p->loc);
  v3 fpara = inplane_unit_perp(fnorm, mdir);
  v3 fparash = cross(fnorm, fpara);
  double veli = veloc_at(d, p->cell, p->type, p->loc);
  double velo = veloc_at(d, other, p->type, p->loc);
  double sini = dot(fpara, mdir);
  double sino = (velo / veli) * sini;
  int transfer; double coso;
  if (sino >= 1.0) { transfer = 0; sino = sini; coso = -1.0 * dot(fnorm, mdir); }
  else { transfer = 1; coso = sqrt(1.0 - (sino * sino)); }
  v3 outdir = add(scal(fpara, sino), scal(fnorm, coso));
  double polout = 0;
  if (p->type != R3D_RAY_P) {
    v3 pdomi = dir_of_motion(p->type, p->th, p->ph, p->pol);
    v3 svbasei = cross(fparash, mdir);
    v3 svbaseo = cross(fparash, outdir);
    double shcomi = dot(pdomi, fparash);
    double svcomi = dot(pdomi, svbasei);
    v3 pdomo = add(scal(fparash, shcomi), scal(svbaseo, svcomi));
    double pol_v = dot(pdomo, xyz_thetahat(outdir));
    double pol_h = dot(pdomo, xyz_phihat(outdir));
    polout = atan2(pol_h, pol_v);
  }
  p->th = xyz_theta(outdir); p->ph = xyz_phi(outdir);
  p->pol = polout;
  if (transfer) p->cell = other;
}
/* CellFace::VelocityJump (media_cellface.cpp:83-99) */
static double velocity_jump(const r3d_model_desc *d, uint32_t cell, uint32_t other, v3 loc) {
  double v1 = veloc_at(d, cell, 0, loc), v2 = veloc_at(d, other, 0, loc);
  double dvp = fabs(2 * (v2 - v1) / (v2 + v1));
  v1 = veloc_at(d, cell, 1, loc); v2 = veloc_at(d, other, 1, loc);
  double dvs = fabs(2 * (v2 - v1) / (v2 + v1));
  return (dvp > dvs) ? dvp : dvs;
}

typedef struct { double *energies; uint64_t *counts; uint64_t *counters; } accum_t;

static void report_collected(const r3d_model_desc *d, phonon_t *p, accum_t *A) {   /* dataout.cpp:545-568 */
  double vel = veloc_at(d, p->cell, p->type, p->loc);
  for (uint32_t s = 0; s < d->n_seis; s++) {
    uint32_t bin; double e[4];
    if (seis_catch(d->seis + (size_t)s * R3D_SEIS_NPARAM, d->bin_dt, d->n_bins, p->time, p->loc, p->th, p->ph,
                   p->pol, p->type, p->amp, vel, &bin, e)) {
      size_t b = (size_t)s * d->n_bins + bin;
      if (A->energies) {
        A->energies[b * 5 + 0] += e[0]; A->energies[b * 5 + 1] += e[1]; A->energies[b * 5 + 2] += e[2];
        A->energies[b * 5 + 3 + p->type] += e[3];
      }
      if (A->counts) A->counts[b * 2 + p->type] += 1;
      p->catches++;
    }
  }
}

/* returns fate (R3D_FATE_* | reason<<8) */
static uint32_t propagate(const r3d_model_desc *d, phonon_t *p, rng_t *g, accum_t *A) {
  for (;;) {
    p->iters++;
    if (p->time > d->ttl) return R3D_FATE_TIMEOUT;
    if ((p->moves % 128) == 127) {                                   /* phonons.cpp:554-584 */
      int why = -1;
      if (isnan(p->pathlen)) why = R3D_INV_PATH_NAN;
      else if (isnan(p->time)) why = R3D_INV_TIME_NAN;
      else if (p->pathlen < 0) why = R3D_INV_PATH_NEGATIVE;
      else if ((p->time < 0) || (p->recent < 0)) why = R3D_INV_TIME_NEGATIVE;
      else if (p->recent == 0) why = R3D_INV_STUCK;
      else if (p->recent < d->slow_concern) why = R3D_INV_SLOW;
      else if (p->moves > d->loop_concern) why = R3D_INV_LOOP_EXCEED;
      if (why >= 0) return R3D_FATE_INVALID | ((1u << why) << 8);
      p->recent = 0;
    }
    travel_t tr = path_to_boundary(d, p->cell, p->type, p->loc, p->th, p->ph);
    if (tr.len == PINF) return R3D_FATE_TIMEOUT;
    uint32_t scat = d->cell_scat[p->cell];
    /* Scatterer::GetRandomPathLength (scatterers.cpp:297-307) */
    double r = ((double)rng_next(g)) / (R3D_RAND_MAX + 1);
    r = 1.0 - r;
    double scatlen = -log(r) * d->scat_mfp[scat * 2 + p->type];
    if (scatlen < tr.len) {
      tr = advance_length(d, p->cell, p->type, scatlen, p->loc, p->th, p->ph);
      ph_move(p, &tr);
      /* Scatterer::GetRandomScatteredRelativePhonon (scatterers.cpp:318-363) */
      double rth, rph, rpol = 0; int otype;
      if (d->no_deflect) { rth = nudge(d, 0); rph = 0; otype = p->type; }
      else {
        uint32_t conv = cdf_search(d->scat_whole_cdf + (scat * 2 + p->type) * 4, 4, rng_next(g));
        otype = conv & 1;                                             /* PP,PS,SP,SS -> P,S,P,S */
        uint32_t ti = cdf_search(d->scat_cdf + ((size_t)scat * 4 + conv) * d->n_toa, d->n_toa, rng_next(g));
        if (conv == 3) rpol = d->scat_spol[(size_t)scat * d->n_toa + ti];
        rth = nudge(d, d->toa_theta[ti]); rph = d->toa_phi[ti];
      }
      transform(&p->th, &p->ph, &p->pol, rth, rph, rpol);
      p->type = otype;
      p->scatters++;
      continue;
    }
    ph_move(p, &tr);
    uint8_t fl = d->face_flags[p->cell * d->faces_per_cell + tr.face];
    if (fl & R3D_FACE_COLLECT) report_collected(d, p, A);
    if (fl & R3D_FACE_REFLECT) { refraction_fullrt(d, p, tr.face, g); continue; }
    if (fl & R3D_FACE_ADJOIN) {                                       /* Phonon::Refract, phonons.cpp:225-255 */
      uint32_t other = d->face_other_cell[p->cell * d->faces_per_cell + tr.face];
      if (fl & R3D_FACE_DISCON) refraction_fullrt(d, p, tr.face, g);
      else if (velocity_jump(d, p->cell, other, p->loc) > 0.00001) refraction_bend(d, p, tr.face);
      else p->cell = other;
      continue;
    }
    return R3D_FATE_LOST;
  }
}

/* ShearDislocation::GenerateEventPhonon (events.cpp:111-124) -> PhononSource::GenerateRandomPhonon
 * (sources.cpp:156-170) -> Phonon ctor (phonons.hpp:193-207) */
static void generate(const r3d_model_desc *d, phonon_t *p, rng_t *g) {
  memset(p, 0, sizeof *p);
  uint32_t rt = cdf_search(d->src_whole_cdf, 3, rng_next(g));
  uint32_t ti = cdf_search(d->src_cdf + (size_t)rt * d->n_toa, d->n_toa, rng_next(g));
  p->amp = 1.0;
  p->th = nudge(d, d->toa_theta[ti]); p->ph = d->toa_phi[ti];
  p->pol = (rt == R3D_RAY_SH) ? R3D_PI * 0.5 : 0.0;
  p->type = (rt == R3D_RAY_P) ? R3D_RAY_P : R3D_RAY_S;
  p->loc = V(d->src_loc[0], d->src_loc[1], d->src_loc[2]);
  p->cell = d->src_cell;
}

static void run_range(const r3d_model_desc *d, uint64_t first, uint64_t n, uint64_t seed, accum_t *A,
                      r3d_phonon_final *finals) {
  for (uint64_t i = 0; i < n; i++) {
    rng_t g = {seed, first + i, 0};
    phonon_t p;
    generate(d, &p, &g);
    uint32_t fate = propagate(d, &p, &g, A);
    if (A->counters) {
      switch (fate & 0xFF) {
      case R3D_FATE_LOST: A->counters[R3D_CNT_LOST]++; break;
      case R3D_FATE_TIMEOUT: A->counters[R3D_CNT_TIMEOUT]++; break;
      default: A->counters[R3D_CNT_INVALID]++; A->counters[7] |= (fate >> 8); break;
      }
      A->counters[R3D_CNT_EVENTS] += p.iters;
      A->counters[R3D_CNT_CATCHES] += p.catches;
      A->counters[R3D_CNT_SCATTERS] += p.scatters;
      A->counters[6] += 1;
    }
    if (finals) {
      r3d_phonon_final *f = finals + i;
      f->time = p.time; f->pathlen = p.pathlen; f->amp = p.amp;
      f->loc[0] = p.loc.x; f->loc[1] = p.loc.y; f->loc[2] = p.loc.z;
      f->theta = p.th; f->phi = p.ph; f->pol = p.pol;
      f->moves = p.moves; f->cell = p.cell; f->type = p.type; f->fate = fate;
      f->draws = g.ordinal; f->catches = p.catches; f->scatters = p.scatters; f->iters = p.iters;
    }
  }
}

typedef struct {
  const r3d_model_desc *d; uint64_t first, n, seed; r3d_phonon_final *finals;
  double *e; uint64_t *c; uint64_t k[R3D_NCOUNTERS];
} job_t;
static void *job_main(void *arg) {
  job_t *j = (job_t *)arg;
  accum_t A = {j->e, j->c, j->k};
  run_range(j->d, j->first, j->n, j->seed, &A, j->finals);
  return NULL;
}

int r3d_oracle_run(const r3d_model_desc *d, uint64_t first, uint64_t n, uint64_t seed,
                   double *energies, uint64_t *counts, uint64_t *counters,
                   r3d_phonon_final *finals, int nthreads) {
  size_t nb = (size_t)d->n_seis * d->n_bins;
  if (nthreads <= 1) {
    accum_t A = {energies, counts, counters};
    run_range(d, first, n, seed, &A, finals);
    return 0;
  }
  /* one contiguous index range and one private set of bins per thread, summed at the end
   * (the reference's own way of combining runs: vis/seisplot/combine.m:26-33) */
  job_t *jobs = (job_t *)calloc((size_t)nthreads, sizeof(job_t));
  pthread_t *tid = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
  if (!jobs || !tid) { free(jobs); free(tid); return 1; }
  int fail = 0, started = 0;
  for (int t = 0; t < nthreads; t++) {
    uint64_t lo = n / (uint64_t)nthreads * (uint64_t)t + (n % (uint64_t)nthreads) * (uint64_t)t / (uint64_t)nthreads;
    uint64_t hi = n / (uint64_t)nthreads * (uint64_t)(t + 1) + (n % (uint64_t)nthreads) * (uint64_t)(t + 1) / (uint64_t)nthreads;
    job_t *j = jobs + t;
    j->d = d; j->first = first + lo; j->n = hi - lo; j->seed = seed; j->finals = finals ? finals + lo : NULL;
    j->e = energies ? (double *)calloc(nb * 5 + 1, sizeof(double)) : NULL;
    j->c = counts ? (uint64_t *)calloc(nb * 2 + 1, sizeof(uint64_t)) : NULL;
    if ((energies && !j->e) || (counts && !j->c)) { fail = 1; free(j->e); free(j->c); break; }
    if (pthread_create(&tid[t], NULL, job_main, j) != 0) { fail = 1; free(j->e); free(j->c); break; }
    started++;
  }
  for (int t = 0; t < started; t++) {
    job_t *j = jobs + t;
    pthread_join(tid[t], NULL);
    if (!fail) {
      if (j->e) for (size_t i = 0; i < nb * 5; i++) energies[i] += j->e[i];
      if (j->c) for (size_t i = 0; i < nb * 2; i++) counts[i] += j->c[i];
      if (counters) for (int i = 0; i < R3D_NCOUNTERS; i++) { if (i == 7) counters[i] |= j->k[i]; else counters[i] += j->k[i]; }
    }
    free(j->e); free(j->c);
  }
  free(jobs); free(tid);
  return fail;
}

/* ---- sub-kernel entry points (layouts: include/r3d_gpu.h r3d_test_*) ------ */
void r3d_oracle_cdf_search(const double *cdf, uint32_t n_cdf, const uint32_t *k, uint32_t n, uint32_t *out) {
  for (uint32_t i = 0; i < n; i++) out[i] = cdf_search(cdf, n_cdf, k[i]);
}
static void put_travel(const travel_t *t, double *o) {
  o[0] = t->len; o[1] = t->time; o[2] = t->loc.x; o[3] = t->loc.y; o[4] = t->loc.z;
  o[5] = t->th; o[6] = t->ph; o[7] = t->atten; o[8] = (double)t->face;
}
void r3d_oracle_path_to_boundary(const r3d_model_desc *d, const double *in, uint32_t n, double *out) {
  for (uint32_t i = 0; i < n; i++) {
    const double *x = in + 7 * i;
    travel_t t = path_to_boundary(d, (uint32_t)x[0], (int)x[1], V(x[2], x[3], x[4]), x[5], x[6]);
    put_travel(&t, out + 9 * i);
  }
}
void r3d_oracle_advance(const r3d_model_desc *d, const double *in, uint32_t n, double *out) {
  for (uint32_t i = 0; i < n; i++) {
    const double *x = in + 8 * i;
    travel_t t = advance_length(d, (uint32_t)x[0], (int)x[1], x[7], V(x[2], x[3], x[4]), x[5], x[6]);
    put_travel(&t, out + 9 * i);
  }
}
void r3d_oracle_transform(const double *in, uint32_t n, double *out) {
  for (uint32_t i = 0; i < n; i++) {
    const double *x = in + 6 * i;
    double th = x[0], ph = x[1], pol = x[2];
    transform(&th, &ph, &pol, x[3], x[4], x[5]);
    out[3 * i] = th; out[3 * i + 1] = ph; out[3 * i + 2] = pol;
  }
}
void r3d_oracle_rtcoef(const double *in, uint32_t n, double *out) {
  for (uint32_t i = 0; i < n; i++) {
    const double *x = in + 15 * i;
    double *o = out + 13 * i;
    rtcoef_t rt;
    rt_init(&rt, V(x[0], x[1], x[2]), V(x[3], x[4], x[5]));
    rt.densR = x[6]; rt.velR[0] = x[7]; rt.velR[1] = x[8];
    rt.densT = x[9]; rt.velT[0] = x[10]; rt.velT[1] = x[11];
    rt.notransmit = x[13] != 0;
    rt_get_coefs(&rt, (int)x[12]);
    rt_choose(&rt, (uint32_t)x[14]);
    v3 od = rt_chosen_dir(&rt), pd = rt_chosen_pdom(&rt);
    for (int k = 0; k < 6; k++) o[k] = rt.prob[k];
    o[6] = rt.choice; o[7] = od.x; o[8] = od.y; o[9] = od.z; o[10] = pd.x; o[11] = pd.y; o[12] = pd.z;
  }
}
void r3d_oracle_catch(double bin_dt, uint32_t n_bins, const double *in, uint32_t n, double *out) {
  for (uint32_t i = 0; i < n; i++) {
    const double *x = in + 28 * i;
    double *o = out + 6 * i;
    uint32_t bin = 0; double e[4] = {0, 0, 0, 0};
    int c = seis_catch(x, bin_dt, n_bins, x[18], V(x[19], x[20], x[21]), x[22], x[23], x[24], (int)x[25], x[26], x[27], &bin, e);
    o[0] = c; o[1] = c ? (double)bin : -1.0; o[2] = e[0]; o[3] = e[1]; o[4] = e[2]; o[5] = e[3];
  }
}


/* ===========================================================================
 * Scatterer tables (scatterers.cpp:134-220, scatparams.cpp:75-194)
 * ======================================================================== */
static double psato(const r3d_scatter_params *P, double m) {            /* scatparams.cpp:181-194 */
  const double pi32 = pow(3.14159265358979323846, 1.5);
  const double numer = (8. * pi32 * P->eps * P->eps * P->a * P->a * P->a) * tgamma(P->kappa + 1.5) / tgamma(P->kappa);
  const double denom = pow((1. + P->a * P->a * m * m), (P->kappa + 1.5));
  return numer / denom;
}
void r3d_oracle_build_scatterer_tables(const r3d_scatter_params *P, const double *toa_theta, const double *toa_phi,
                                       uint32_t n_toa, double *cdf, double *spol, double *whole_cdf, double *mfp) {
  const size_t n = n_toa;
  const double pi = 3.14159265358979323846;
  for (size_t k = 0; k < n; k++) {
    /* XSATO, scatparams.cpp:136-160 */
    const double th = toa_theta[k], ph = toa_phi[k];
    const double gam0 = P->gam0, nu = P->nu, el = P->el;
    const double gam2x = gam0 * gam0;
    const double cpsi = cos(th), c2psi = cos(2. * th), spsi = sin(th), czeta = cos(ph), szeta = sin(ph), spsi2 = spsi * spsi;
    const double xpp = (1. / gam2x) * (nu * (-1. + cpsi + (2. / gam2x) * spsi2) - 2. + (4. / gam2x) * spsi2);
    const double xps = -spsi * (nu * (1. - (2. / gam0) * cpsi) - (4. / gam0) * cpsi);
    const double xsp = (1. / gam2x) * spsi * czeta * (nu * (1. - (2. / gam0) * cpsi) - (4. / gam0) * cpsi);
    const double xss_psi = czeta * (nu * (cpsi - c2psi) - 2. * c2psi);
    const double xss_zeta = szeta * (nu * (cpsi - 1.) + 2. * cpsi);
    /* GSATO, scatparams.cpp:75-118 */
    const double pi4 = 4. * pi, el4 = pow(el, 4), gam2 = pow(gam0, 2), psi = th;
    double arg = (2. * el / gam0) * sin(psi / 2.);
    double gpp = (el4 / pi4) * (xpp * xpp) * psato(P, arg);
    if (gpp < 1.e-30) gpp = 0.;
    arg = (el / gam0) * sqrt(1. + gam2 - 2. * gam0 * cos(psi));
    double gps = (1. / gam0) * (el4 / pi4) * (xps * xps) * psato(P, arg);
    if (gps < 1.e-30) gps = 0.;
    double gsp = gam0 * (el4 / pi4) * (xsp * xsp) * psato(P, arg);
    if (gsp < 1.e-30) gsp = 0.;
    arg = 2. * el * sin(psi / 2.);
    double gss = (el4 / pi4) * (xss_psi * xss_psi + xss_zeta * xss_zeta) * psato(P, arg);
    if (gss < 1.e-30) gss = 0.;
    cdf[k] = gpp; cdf[n + k] = gps; cdf[2 * n + k] = gsp; cdf[3 * n + k] = gss;
    spol[k] = atan2(xss_zeta, xss_psi);
  }
  for (int t = 0; t < 4; t++) {            /* ProbDist::Integrate, probability.cpp:21-35 */
    double *c = cdf + (size_t)t * n;
    for (size_t i = 1; i < n; i++) c[i] += c[i - 1];
  }
  /* PopulateWholeProbs (scatterers.cpp:170-183), integrated */
  whole_cdf[0] = cdf[n - 1]; whole_cdf[1] = whole_cdf[0] + cdf[2 * n - 1]; whole_cdf[2] = whole_cdf[1] + 0.0; whole_cdf[3] = whole_cdf[2] + 0.0;
  whole_cdf[4] = 0.0; whole_cdf[5] = 0.0; whole_cdf[6] = 0.0 + cdf[3 * n - 1]; whole_cdf[7] = whole_cdf[6] + cdf[4 * n - 1];
  /* ComputeMFPs, scatterers.cpp:195-220 */
  mfp[0] = 1.0 / (whole_cdf[3] / (double)n_toa);
  mfp[1] = 1.0 / (whole_cdf[7] / (double)n_toa);
}
