// r3d_gpu.cu -- propagate kernels and the C ABI of include/r3d_gpu.h (sm_100a only).
//
// One phonon per thread.  A warp keeps all 32 lanes busy by refilling dead lanes from a
// per-launch work counter (chunked, one atomic per warp per 128 phonons); each loop
// iteration is one iteration of the reference's Propagate loop (phonons.cpp:542) for every
// live lane, arranged in three phases so that the expensive, latency-bound CDF search is
// executed once per iteration for every lane that needs one -- whether it is a freshly
// generated source phonon (sources.cpp:156-170) or a scatter draw (scatterers.cpp:318-363):
//   A  refill / time-out + validity / distance to boundary / path-length draw / boundary work
//   B  CDF search + take-off-angle fetch            (lanes with a draw request)
//   C  new-phonon init, or Phonon::Transform
// There is no CPU fallback anywhere in this file.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>
#include <algorithm>
#include "r3d_gpu.h"
#include "r3d_device.cuh"

using namespace r3d;

#define R3D_THREADS 128
#define R3D_CHUNK 128ull
#define FULL 0xffffffffu

namespace {

thread_local std::string g_err;
int fail(int code, const std::string &msg) { g_err = msg; return code; }
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e_ = (call);                                                                       \
    if (e_ != cudaSuccess)                                                                         \
      return fail(e_ == cudaErrorMemoryAllocation ? R3D_ENOMEM : R3D_ECUDA,                        \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                             \
  } while (0)

// ---------------------------------------------------------------------------
// phonon state (phonons.hpp:69-126), register resident
// ---------------------------------------------------------------------------
struct Phonon {
  double time, pathlen, recent, amp;
  v3 loc;
  double th, ph, pol;
  uint32_t moves, cell;
  int type;
};

R3D_DEV void move(Phonon &p, const Travel &t) {   // Phonon::Move, phonons.cpp:62-70
  p.pathlen += t.len; p.time += t.time; p.recent += t.time;
  p.loc = t.loc; p.th = t.th; p.ph = t.ph;
  p.amp *= t.atten; p.moves += 1;
}

// Phonon::Refraction_FullRT (phonons.cpp:429-476) + CellFace::GetRTBasis (media_cellface.cpp:122-149)
template <class Cell>
R3D_DEV void refraction_fullrt(const DevModel &M, const double *cells, Phonon &p, int face, bool adjoin, uint32_t other, Rng &g) {
  const double *c = cells + (size_t)p.cell * M.cell_nparam;
  RTCoef rt;
  rt.init(Cell::normal(c, face, p.loc), from_thph(p.th, p.ph));
  rt.densR = Cell::dens(c, p.loc);
  rt.velR[0] = Cell::veloc(c, 0, p.loc);
  rt.velR[1] = Cell::veloc(c, 1, p.loc);
  if (adjoin) {
    const double *o = cells + (size_t)other * M.cell_nparam;
    rt.densT = Cell::dens(o, p.loc);
    rt.velT[0] = Cell::veloc(o, 0, p.loc);
    rt.velT[1] = Cell::veloc(o, 1, p.loc);
  } else {                                    // free surface
    rt.densT = 0.0; rt.velT[0] = 1e-12; rt.velT[1] = 1e-12; rt.notransmit = true;
  }
  int intype = R3D_RAY_P;
  if (p.type == R3D_RAY_S) intype = rt.choose_spol(dir_of_motion(p.type, p.th, p.ph, p.pol), g.next());
  rt.get_coefs(intype);
  rt.choose(g.next());
  const bool reflected = (rt.choice == R_P || rt.choice == R_SV || rt.choice == R_SH);
  v3 outdir = rt.chosen_ray_dir();
  p.type = (rt.choice == R_P || rt.choice == T_P) ? R3D_RAY_P : R3D_RAY_S;
  p.th = xyz_theta(outdir); p.ph = xyz_phi(outdir);
  if (p.type == R3D_RAY_S) {
    v3 pdomo = rt.chosen_pdom();
    p.pol = atan2(dot(pdomo, thph_phihat(p.ph)), dot(pdomo, thph_thetahat(p.th, p.ph)));
  }
  if (!reflected) p.cell = other;
}

// Phonon::Refraction_Bend (phonons.cpp:311-405)
template <class Cell>
R3D_DEV void refraction_bend(const DevModel &M, const double *cells, Phonon &p, int face, uint32_t other) {
  const double *c = cells + (size_t)p.cell * M.cell_nparam;
  const double *o = cells + (size_t)other * M.cell_nparam;
  v3 mdir = from_thph(p.th, p.ph);
  v3 fnorm = Cell::normal(c, face, p.loc);
  v3 fpara = inplane_unit_perp(fnorm, mdir);
  v3 fparash = cross(fnorm, fpara);
  double veli = Cell::veloc(c, p.type, p.loc), velo = Cell::veloc(o, p.type, p.loc);
  double sini = dot(fpara, mdir);
  double sino = (velo / veli) * sini;
  bool transfer; double coso;
  if (sino >= 1.0) { transfer = false; sino = sini; coso = -1.0 * dot(fnorm, mdir); }
  else { transfer = true; coso = sqrt(1.0 - (sino * sino)); }
  v3 outdir = add(scal(fpara, sino), scal(fnorm, coso));
  double polout = 0;
  if (p.type != R3D_RAY_P) {
    v3 pdomi = dir_of_motion(p.type, p.th, p.ph, p.pol);
    v3 svbasei = cross(fparash, mdir), svbaseo = cross(fparash, outdir);
    double shcomi = dot(pdomi, fparash), svcomi = dot(pdomi, svbasei);
    v3 pdomo = add(scal(fparash, shcomi), scal(svbaseo, svcomi));
    polout = atan2(dot(pdomo, xyz_phihat(outdir)), dot(pdomo, xyz_thetahat(outdir)));
  }
  p.th = xyz_theta(outdir); p.ph = xyz_phi(outdir);
  p.pol = polout;
  if (transfer) p.cell = other;
}

// CellFace::VelocityJump (media_cellface.cpp:83-99)
template <class Cell>
R3D_DEV double velocity_jump(const DevModel &M, const double *cells, uint32_t cell, uint32_t other, v3 loc) {
  const double *c = cells + (size_t)cell * M.cell_nparam, *o = cells + (size_t)other * M.cell_nparam;
  double v1 = Cell::veloc(c, 0, loc), v2 = Cell::veloc(o, 0, loc);
  double dvp = fabs(2 * (v2 - v1) / (v2 + v1));
  v1 = Cell::veloc(c, 1, loc); v2 = Cell::veloc(o, 1, loc);
  double dvs = fabs(2 * (v2 - v1) / (v2 + v1));
  return (dvp > dvs) ? dvp : dvs;
}

// DataReporter::ReportPhononCollected (dataout.cpp:545-568): every seismometer is pass-through
// (dataout.cpp:50), so all of them are tested.  A conservative squared-distance pre-filter on a
// shared-memory (x,y,z,r_out^2) record skips the exact test for seismometers that cannot catch.
template <class Cell>
R3D_DEV uint32_t collect(const DevModel &M, const double *cells, const double4 *sph, const Phonon &p) {
  const double *c = cells + (size_t)p.cell * M.cell_nparam;
  const double vel = Cell::veloc(c, p.type, p.loc);
  const v3 dir = from_thph(p.th, p.ph);
  const v3 dopm = dir_of_motion(p.type, p.th, p.ph, p.pol);
  uint32_t n = 0;
  for (uint32_t s = 0; s < M.n_seis; s++) {
    double4 q = sph[s];
    double dx = q.x - p.loc.x, dy = q.y - p.loc.y, dz = q.z - p.loc.z;
    if (dx * dx + dy * dy + dz * dz > q.w) continue;
    uint32_t bin; double e[4];
    if (seis_catch(M.seis + (size_t)s * R3D_SEIS_NPARAM, M.bin_dt, M.n_bins, p.time, p.loc, dir, dopm, p.type, p.amp, vel, bin, e)) {
      size_t b = (size_t)s * M.n_bins + bin;
      atomicAdd(M.energies + b * 5 + 0, e[0]);
      atomicAdd(M.energies + b * 5 + 1, e[1]);
      atomicAdd(M.energies + b * 5 + 2, e[2]);
      atomicAdd(M.energies + b * 5 + 3 + p.type, e[3]);
      atomicAdd(M.counts + b * 2 + p.type, 1ull);
      n++;
    }
  }
  return n;
}

R3D_DEV void write_final(r3d_phonon_final *f, const Phonon &p, uint32_t fate, const Rng &g, uint32_t catches, uint32_t scatters, uint32_t iters) {
  f->time = p.time; f->pathlen = p.pathlen; f->amp = p.amp;
  f->loc[0] = p.loc.x; f->loc[1] = p.loc.y; f->loc[2] = p.loc.z;
  f->theta = p.th; f->phi = p.ph; f->pol = p.pol;
  f->moves = p.moves; f->cell = p.cell; f->type = (uint32_t)p.type; f->fate = fate;
  f->draws = g.ordinal; f->catches = catches; f->scatters = scatters; f->iters = iters;
}

// ---------------------------------------------------------------------------
// the propagate kernel
// ---------------------------------------------------------------------------
template <class Cell, bool TRACE>
__global__ void __launch_bounds__(R3D_THREADS)
propagate_kernel(const DevModel M, unsigned long long first, unsigned long long n, unsigned long long seed,
                 int cells_in_smem, r3d_phonon_final *finals) {
  extern __shared__ double4 smem4[];
  double4 *sph = smem4;                                       // [n_seis]
  double *scells = reinterpret_cast<double *>(smem4 + M.n_seis);
  for (uint32_t i = threadIdx.x; i < M.n_seis; i += blockDim.x) sph[i] = M.seis_sphere[i];
  if (cells_in_smem)
    for (uint32_t i = threadIdx.x; i < M.n_cells * M.cell_nparam; i += blockDim.x) scells[i] = M.cell_params[i];
  __syncthreads();
  const double *cells = cells_in_smem ? scells : M.cell_params;

  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt_mask = (1u << lane) - 1u;
  Phonon p;
  Rng g;
  bool alive = false, done = false;
  unsigned long long idx = 0, wnext = 0, wend = 0;
  uint32_t ph_catches = 0, ph_scatters = 0, ph_iters = 0;
  // per-thread tallies (dataout.cpp:591-617)
  unsigned long long n_lost = 0, n_timeout = 0, n_invalid = 0, n_events = 0, n_catches = 0, n_scatters = 0, n_phonons = 0;
  uint32_t diag = 0;
  g.init(seed, 0);
  p.time = p.pathlen = p.recent = p.amp = 0; p.loc = V(0, 0, 0); p.th = p.ph = p.pol = 0; p.moves = 0; p.cell = 0; p.type = 0;

  for (;;) {
    int req = 0;                 // 0 none, 1 source take-off angle, 2 scatter angle
    const double *rcdf = nullptr; const uint32_t *rguide = nullptr; uint32_t rk = 0, conv = 0;

    // ---- phase A.0: refill dead lanes --------------------------------------
    const bool need = !alive && !done;
    const unsigned needmask = __ballot_sync(FULL, need);
    if (needmask) {
      if (wnext >= wend) {                                   // warp-uniform
        unsigned long long base = 0;
        if (lane == 0) base = atomicAdd(M.next_phonon, R3D_CHUNK);
        base = __shfl_sync(FULL, base, 0);
        wnext = (base < n) ? base : n;
        wend = (base + R3D_CHUNK < n) ? base + R3D_CHUNK : n;
      }
      if (need) {
        unsigned long long cand = wnext + __popc(needmask & lt_mask);
        if (cand < wend) {
          // ShearDislocation::GenerateEventPhonon (events.cpp:111-124)
          idx = first + cand;
          g.init(seed, idx);
          uint32_t rt3 = cdf_search_small(M.src_whole, 3, g.next());
          rcdf = M.src_cdf + (size_t)rt3 * M.n_toa;
          rguide = M.src_guide + (size_t)rt3 * M.guide_stride;
          rk = g.next();
          req = 1;
          p.time = 0; p.pathlen = 0; p.recent = 0; p.moves = 0; p.amp = 1.0;       // phonons.hpp:193-207
          p.loc = V(M.src_loc[0], M.src_loc[1], M.src_loc[2]);
          p.cell = M.src_cell;
          p.pol = (rt3 == R3D_RAY_SH) ? kPi * 0.5 : 0.0;
          p.type = (rt3 == R3D_RAY_P) ? R3D_RAY_P : R3D_RAY_S;
          ph_catches = ph_scatters = ph_iters = 0;
          n_phonons++;
        } else if (wend >= n) {
          done = true;
        }
      }
      unsigned long long adv = wnext + __popc(needmask);
      wnext = (adv < wend) ? adv : wend;
    }
    if (__all_sync(FULL, done && !alive)) break;

    // ---- phase A: one Propagate-loop iteration (phonons.cpp:542-679) ---------
    if (alive) {
      uint32_t fate = 0;
      n_events++; ph_iters++;
      if (p.time > M.ttl) fate = R3D_FATE_TIMEOUT;
      else if ((p.moves % 128u) == 127u) {                   // phonons.cpp:554-584
        int why = -1;
        if (isnan(p.pathlen)) why = R3D_INV_PATH_NAN;
        else if (isnan(p.time)) why = R3D_INV_TIME_NAN;
        else if (p.pathlen < 0) why = R3D_INV_PATH_NEGATIVE;
        else if ((p.time < 0) || (p.recent < 0)) why = R3D_INV_TIME_NEGATIVE;
        else if (p.recent == 0) why = R3D_INV_STUCK;
        else if (p.recent < M.slow_concern) why = R3D_INV_SLOW;
        else if (p.moves > M.loop_concern) why = R3D_INV_LOOP_EXCEED;
        if (why >= 0) fate = R3D_FATE_INVALID | ((1u << why) << 8);
        else p.recent = 0;
      }
      if (!fate) {
        const double *c = cells + (size_t)p.cell * M.cell_nparam;
        typename Cell::Path P;
        const double edgelen = Cell::path(M, c, p.type, p.loc, p.th, p.ph, P);
        if (edgelen == pinf()) fate = R3D_FATE_TIMEOUT;       // phonons.cpp:595-598
        else {
          const uint32_t scat = __ldg(M.cell_scat + p.cell);
          // Scatterer::GetRandomPathLength (scatterers.cpp:297-307)
          double r = 1.0 - ((double)g.next()) / (kRandMax + 1);
          const double scatlen = -log(r) * __ldg(M.scat_mfp + scat * 2 + p.type);
          if (scatlen < edgelen) {
            Travel tr = Cell::advance(M, c, p.type, scatlen, p.loc, p.th, p.ph, P);
            move(p, tr);
            // Scatterer::GetRandomScatteredRelativePhonon (scatterers.cpp:318-363)
            if (M.no_deflect) {
              transform(p.th, p.ph, p.pol, M.min_theta, 0.0, 0.0);
              n_scatters++; ph_scatters++;
            } else {
              conv = cdf_search_small(M.scat_whole + (scat * 2 + p.type) * 4, 4, g.next());
              rcdf = M.scat_cdf + ((size_t)scat * 4 + conv) * M.n_toa;
              rguide = M.scat_guide + ((size_t)scat * 4 + conv) * M.guide_stride;
              rk = g.next();
              req = 2;
            }
          } else {
            Travel tr = Cell::advance(M, c, p.type, edgelen, p.loc, p.th, p.ph, P);
            move(p, tr);
            const uint32_t fi = p.cell * M.faces_per_cell + P.face;
            const uint32_t fl = __ldg(M.face_flags + fi);
            const uint32_t other = __ldg(M.face_other + fi);
            if (fl & R3D_FACE_COLLECT) {
              uint32_t k = collect<Cell>(M, cells, sph, p);
              n_catches += k; ph_catches += k;
            }
            if (fl & R3D_FACE_REFLECT) refraction_fullrt<Cell>(M, cells, p, P.face, (fl & R3D_FACE_ADJOIN) != 0, other, g);
            else if (fl & R3D_FACE_ADJOIN) {                 // Phonon::Refract, phonons.cpp:225-255
              if (fl & R3D_FACE_DISCON) refraction_fullrt<Cell>(M, cells, p, P.face, true, other, g);
              else if (velocity_jump<Cell>(M, cells, p.cell, other, p.loc) > 0.00001) refraction_bend<Cell>(M, cells, p, P.face, other);
              else p.cell = other;                           // Refraction_Continuous
            } else fate = R3D_FATE_LOST;
          }
        }
      }
      if (fate) {
        alive = false;
        switch (fate & 0xFF) {
          case R3D_FATE_LOST: n_lost++; break;
          case R3D_FATE_TIMEOUT: n_timeout++; break;
          default: n_invalid++; diag |= (fate >> 8); break;
        }
        if (TRACE) write_final(finals + (idx - first), p, fate, g, ph_catches, ph_scatters, ph_iters);
      }
    }

    // ---- phase B + C: CDF search, take-off angle, new direction ------------------
    if (req) {
      const uint32_t ti = cdf_search(rcdf, M.n_toa, rguide, M.guide_shift, rk);
      const double2 t = __ldg(M.toa + ti);
      if (req == 1) { p.th = t.x; p.ph = t.y; alive = true; }
      else {
        const uint32_t scat = __ldg(M.cell_scat + p.cell);
        const double rpol = (conv == 3) ? __ldg(M.scat_spol + (size_t)scat * M.n_toa + ti) : 0.0;
        transform(p.th, p.ph, p.pol, t.x, t.y, rpol);
        p.type = (int)(conv & 1u);                            // PP,PS,SP,SS -> P,S,P,S
        n_scatters++; ph_scatters++;
      }
    }
  }

  // ---- tallies: warp reduce, one atomic per warp and counter ----------------------
  unsigned long long t[7] = {n_lost, n_timeout, n_invalid, n_events, n_catches, n_scatters, n_phonons};
#pragma unroll
  for (int k = 0; k < 7; k++) {
    unsigned long long v = t[k];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(FULL, v, o);
    if (lane == 0 && v) atomicAdd(M.counters + k, v);
  }
  diag = __reduce_or_sync(FULL, diag);
  if (lane == 0 && diag) atomicOr(M.counters + 7, (unsigned long long)diag);
}

// ---------------------------------------------------------------------------
// model-preparation kernels
// ---------------------------------------------------------------------------
__global__ void pack_toa_kernel(const double *th, const double *ph, double2 *out, uint32_t n, double mn, double mx) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double t = th[i];
  if (t < mn) t = mn;                    // Phonon::nudge_if_singular (phonons.hpp:335-344), applied where the
  if (t > mx) t = mx;                    // reference constructs a Phonon from a TOA entry
  out[i] = make_double2(t, ph[i]);
}
// guide[j] = lower_bound of r(k = min(j << shift, RAND_MAX)) for every table
__global__ void build_guide_kernel(const double *cdf, uint32_t n_toa, uint32_t n_tables, uint32_t shift, uint32_t stride, uint32_t *guide) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned long long)n_tables * stride) return;
  uint32_t t = (uint32_t)(i / stride), j = (uint32_t)(i % stride);
  unsigned long long k = (unsigned long long)j << shift;
  if (k > 2147483647ull) k = 2147483647ull;
  guide[i] = cdf_search_plain(cdf + (size_t)t * n_toa, n_toa, (uint32_t)k);
}
__global__ void check_monotone_kernel(const double *cdf, uint32_t n_toa, uint32_t n_tables, int *bad) {
  unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (unsigned long long)n_tables * n_toa) return;
  if ((i % n_toa) == 0) { if (!(cdf[i] >= 0.0)) *bad = 1; return; }
  if (!(cdf[i] >= cdf[i - 1])) *bad = 1;
}
__global__ void seis_sphere_kernel(const double *seis, uint32_t n, double4 *out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double *s = seis + (size_t)i * R3D_SEIS_NPARAM;
  double ro = fmax(s[14], s[15]);
  out[i] = make_double4(s[0], s[1], s[2], ro * ro * (1.0 + 1e-12));   // d^2 > w  =>  sqrt(d^2) > r_out for both types
}

// ---------------------------------------------------------------------------
// sub-kernel hooks (r3d_test_*)
// ---------------------------------------------------------------------------
__global__ void test_cdf_kernel(const double *cdf, uint32_t n_cdf, const uint32_t *guide, uint32_t shift, const uint32_t *k, uint32_t n, uint32_t *out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = (n_cdf <= 4) ? cdf_search_small(cdf, (int)n_cdf, k[i]) : cdf_search(cdf, n_cdf, guide, shift, k[i]);
}
template <class Cell>
__global__ void test_path_kernel(const DevModel M, const double *in, uint32_t n, double *out, int advance_mode) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double *x = in + (advance_mode ? 8 : 7) * i;
  const double *c = M.cell_params + (size_t)(uint32_t)x[0] * M.cell_nparam;
  int rt = (int)x[1];
  v3 loc = V(x[2], x[3], x[4]);
  typename Cell::Path P;
  double len = Cell::path(M, c, rt, loc, x[5], x[6], P);
  Travel t = Cell::advance(M, c, rt, advance_mode ? x[7] : len, loc, x[5], x[6], P);
  double *o = out + 9 * i;
  o[0] = t.len; o[1] = t.time; o[2] = t.loc.x; o[3] = t.loc.y; o[4] = t.loc.z; o[5] = t.th; o[6] = t.ph; o[7] = t.atten;
  o[8] = advance_mode ? -1.0 : (double)P.face;
}
__global__ void test_transform_kernel(const double *in, uint32_t n, double *out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double *x = in + 6 * i;
  double th = x[0], ph = x[1], pol = x[2];
  transform(th, ph, pol, x[3], x[4], x[5]);
  out[3 * i] = th; out[3 * i + 1] = ph; out[3 * i + 2] = pol;
}
__global__ void test_rtcoef_kernel(const double *in, uint32_t n, double *out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double *x = in + 15 * i;
  double *o = out + 13 * i;
  RTCoef rt;
  rt.init(V(x[0], x[1], x[2]), V(x[3], x[4], x[5]));
  rt.densR = x[6]; rt.velR[0] = x[7]; rt.velR[1] = x[8];
  rt.densT = x[9]; rt.velT[0] = x[10]; rt.velT[1] = x[11];
  rt.notransmit = x[13] != 0;
  rt.get_coefs((int)x[12]);
  rt.choose((uint32_t)x[14]);
  v3 od = rt.chosen_ray_dir(), pd = rt.chosen_pdom();
  for (int k = 0; k < 6; k++) o[k] = rt.prob[k];
  o[6] = rt.choice; o[7] = od.x; o[8] = od.y; o[9] = od.z; o[10] = pd.x; o[11] = pd.y; o[12] = pd.z;
}
__global__ void test_catch_kernel(double bin_dt, uint32_t n_bins, const double *in, uint32_t n, double *out) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double *x = in + 28 * i;
  double *o = out + 6 * i;
  uint32_t bin = 0; double e[4] = {0, 0, 0, 0};
  int type = (int)x[25];
  bool c = seis_catch(x, bin_dt, n_bins, x[18], V(x[19], x[20], x[21]), from_thph(x[22], x[23]),
                      dir_of_motion(type, x[22], x[23], x[24]), type, x[26], x[27], bin, e);
  o[0] = c ? 1.0 : 0.0; o[1] = c ? (double)bin : -1.0; o[2] = e[0]; o[3] = e[1]; o[4] = e[2]; o[5] = e[3];
}

// ---------------------------------------------------------------------------
// host side of the handle
// ---------------------------------------------------------------------------
struct DevState {
  int device = -1;
  cudaStream_t stream = nullptr;
  DevModel M;
  std::vector<void *> allocs;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> timing;   // one pair per r3d_run since the last sync
  int grid = 0;
  size_t smem = 0;
  int cells_in_smem = 0;
};

}  // namespace

struct r3d_handle {
  std::vector<DevState> devs;
  uint32_t cell_kind = 0, n_seis = 0, n_bins = 0;
  unsigned long long launches = 0;
};

namespace {

template <class T>
int dev_alloc(DevState &D, T **p, size_t count) {
  void *q = nullptr;
  CK(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
  D.allocs.push_back(q);
  *p = static_cast<T *>(q);
  return 0;
}
template <class T>
int dev_upload(DevState &D, const T **p, const T *host, size_t count) {
  T *q = nullptr;
  if (int rc = dev_alloc(D, &q, count)) return rc;
  if (count) CK(cudaMemcpyAsync(q, host, count * sizeof(T), cudaMemcpyHostToDevice, D.stream));
  *p = q;
  return 0;
}

typedef void (*propagate_fn)(const DevModel, unsigned long long, unsigned long long, unsigned long long, int, r3d_phonon_final *);
propagate_fn pick_kernel(uint32_t kind, bool trace) {
  switch (kind) {
    case R3D_CELL_CYLINDER: return trace ? propagate_kernel<Cylinder, true> : propagate_kernel<Cylinder, false>;
    case R3D_CELL_SHELL: return trace ? propagate_kernel<Shell, true> : propagate_kernel<Shell, false>;
    default: return trace ? propagate_kernel<Tetra, true> : propagate_kernel<Tetra, false>;
  }
}

int validate(const r3d_model_desc *d) {
  if (!d) return fail(R3D_EINVAL, "null model descriptor");
  if (!d->n_toa || !d->n_cells || !d->n_scat) return fail(R3D_EINVAL, "model needs at least one take-off angle, cell and scatterer");
  if (!d->toa_theta || !d->toa_phi || !d->src_whole_cdf || !d->src_cdf || !d->scat_mfp || !d->scat_whole_cdf ||
      !d->scat_cdf || !d->scat_spol || !d->cell_params || !d->cell_scat || !d->face_flags || !d->face_other_cell ||
      (d->n_seis && !d->seis))
    return fail(R3D_EINVAL, "null array in model descriptor");
  uint32_t np, nf;
  switch (d->cell_kind) {
    case R3D_CELL_CYLINDER: np = R3D_CYL_NPARAM; nf = R3D_CYL_NFACES; break;
    case R3D_CELL_SHELL: np = R3D_SHELL_NPARAM; nf = R3D_SHELL_NFACES; break;
    case R3D_CELL_TETRA: np = R3D_TETRA_NPARAM; nf = R3D_TETRA_NFACES; break;
    default: return fail(R3D_EINVAL, "unknown cell_kind");
  }
  if (d->cell_nparam != np || d->faces_per_cell != nf) return fail(R3D_EINVAL, "cell_nparam / faces_per_cell do not match cell_kind");
  if (d->src_cell >= d->n_cells) return fail(R3D_EINVAL, "src_cell out of range");
  if (!(d->bin_dt > 0) || !d->n_bins) return fail(R3D_EINVAL, "bin_dt and n_bins must be positive");
  for (uint32_t i = 0; i < d->n_cells; i++) {
    if (d->cell_scat[i] >= d->n_scat) return fail(R3D_EINVAL, "cell_scat out of range");
    for (uint32_t f = 0; f < nf; f++)
      if ((d->face_flags[i * nf + f] & R3D_FACE_ADJOIN) && d->face_other_cell[i * nf + f] >= d->n_cells)
        return fail(R3D_EINVAL, "face_other_cell out of range");
    if (d->cell_kind == R3D_CELL_SHELL) {
      const double *c = d->cell_params + (size_t)i * np;
      if (c[0] > 0 || c[1] > 0)        // media.cpp:675 throws for inverted radial velocity profiles
        return fail(R3D_EUNSUPPORTED, "SphereShell: no handler for inverted radial velocity profiles");
    }
  }
  return 0;
}

int build_device(DevState &D, const r3d_model_desc *d, int guide_bits_req) {
  CK(cudaSetDevice(D.device));
  CK(cudaStreamCreateWithFlags(&D.stream, cudaStreamNonBlocking));
  DevModel &M = D.M;
  memset(&M, 0, sizeof M);
  M.freq_hz = d->freq_hz; M.ttl = d->ttl; M.bin_dt = d->bin_dt;
  for (int i = 0; i < 3; i++) { M.earth_center[i] = d->earth_center[i]; M.src_loc[i] = d->src_loc[i]; M.src_whole[i] = d->src_whole_cdf[i]; }
  M.min_theta = d->min_theta; M.max_theta = d->max_theta; M.slow_concern = d->slow_concern;
  M.cyl_radius2 = d->cyl_radius2; M.loop_concern = d->loop_concern;
  M.n_bins = d->n_bins; M.n_toa = d->n_toa; M.src_cell = d->src_cell; M.n_scat = d->n_scat; M.n_cells = d->n_cells; M.n_seis = d->n_seis;
  M.ecs_radial = d->ecs_radial; M.no_deflect = d->no_deflect;
  M.cell_nparam = d->cell_nparam; M.faces_per_cell = d->faces_per_cell;
  const size_t nt = d->n_toa, ns = d->n_scat, nc = d->n_cells, nf = d->faces_per_cell;

  // take-off angles, packed (theta, phi) with the constructor's theta clamp applied
  const double *th = nullptr, *ph = nullptr;
  if (int rc = dev_upload(D, &th, d->toa_theta, nt)) return rc;
  if (int rc = dev_upload(D, &ph, d->toa_phi, nt)) return rc;
  double2 *toa = nullptr;
  if (int rc = dev_alloc(D, &toa, nt)) return rc;
  pack_toa_kernel<<<(unsigned)((nt + 255) / 256), 256, 0, D.stream>>>(th, ph, toa, (uint32_t)nt, d->min_theta, d->max_theta);
  M.toa = toa;

  if (int rc = dev_upload(D, &M.src_cdf, d->src_cdf, 3 * nt)) return rc;
  if (int rc = dev_upload(D, &M.scat_mfp, d->scat_mfp, 2 * ns)) return rc;
  if (int rc = dev_upload(D, &M.scat_whole, d->scat_whole_cdf, 8 * ns)) return rc;
  if (int rc = dev_upload(D, &M.scat_cdf, d->scat_cdf, 4 * ns * nt)) return rc;
  if (int rc = dev_upload(D, &M.scat_spol, d->scat_spol, ns * nt)) return rc;
  if (int rc = dev_upload(D, &M.cell_params, d->cell_params, nc * d->cell_nparam)) return rc;
  if (int rc = dev_upload(D, &M.cell_scat, d->cell_scat, nc)) return rc;
  if (int rc = dev_upload(D, &M.face_flags, d->face_flags, nc * nf)) return rc;
  if (int rc = dev_upload(D, &M.face_other, d->face_other_cell, nc * nf)) return rc;
  if (int rc = dev_upload(D, &M.seis, d->seis, (size_t)d->n_seis * R3D_SEIS_NPARAM)) return rc;
  double4 *sph = nullptr;
  if (int rc = dev_alloc(D, &sph, d->n_seis)) return rc;
  if (d->n_seis) seis_sphere_kernel<<<(d->n_seis + 127) / 128, 128, 0, D.stream>>>(M.seis, d->n_seis, sph);
  M.seis_sphere = sph;

  // guide tables: exact only for non-decreasing CDFs; otherwise fall back to the plain bisection
  int *bad = nullptr;
  if (int rc = dev_alloc(D, &bad, 1)) return rc;
  CK(cudaMemsetAsync(bad, 0, sizeof(int), D.stream));
  check_monotone_kernel<<<(unsigned)((3 * nt + 255) / 256), 256, 0, D.stream>>>(M.src_cdf, (uint32_t)nt, 3, bad);
  check_monotone_kernel<<<(unsigned)((4 * ns * nt + 255) / 256), 256, 0, D.stream>>>(M.scat_cdf, (uint32_t)nt, (uint32_t)(4 * ns), bad);
  int hbad = 0;
  CK(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, D.stream));
  CK(cudaStreamSynchronize(D.stream));
  int bits = guide_bits_req;
  if (bits < 0) {                         // default: about 4 table entries per bucket
    bits = 0;
    while ((1ull << (bits + 2)) < nt && bits < 24) bits++;
  }
  if (bits > 24) bits = 24;
  if (hbad || bits == 0 || nt < 16) {
    M.guide_shift = 32; M.guide_stride = 1;
    uint32_t *g = nullptr;
    if (int rc = dev_alloc(D, &g, 1)) return rc;
    M.src_guide = g; M.scat_guide = g;
  } else {
    M.guide_shift = 31 - bits;
    M.guide_stride = (1u << bits) + 1;
    uint32_t *gs = nullptr, *gc = nullptr;
    if (int rc = dev_alloc(D, &gs, (size_t)3 * M.guide_stride)) return rc;
    if (int rc = dev_alloc(D, &gc, (size_t)4 * ns * M.guide_stride)) return rc;
    unsigned long long tot = 3ull * M.guide_stride;
    build_guide_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, D.stream>>>(M.src_cdf, (uint32_t)nt, 3, M.guide_shift, M.guide_stride, gs);
    tot = 4ull * ns * M.guide_stride;
    build_guide_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, D.stream>>>(M.scat_cdf, (uint32_t)nt, (uint32_t)(4 * ns), M.guide_shift, M.guide_stride, gc);
    M.src_guide = gs; M.scat_guide = gc;
  }

  // accumulators
  const size_t nb = std::max<size_t>((size_t)d->n_seis * d->n_bins, 1);
  if (int rc = dev_alloc(D, &M.energies, nb * R3D_BIN_NF64)) return rc;
  if (int rc = dev_alloc(D, &M.counts, nb * R3D_BIN_NCNT)) return rc;
  if (int rc = dev_alloc(D, &M.counters, (size_t)R3D_NCOUNTERS)) return rc;
  if (int rc = dev_alloc(D, &M.next_phonon, (size_t)1)) return rc;
  CK(cudaMemsetAsync(M.energies, 0, nb * R3D_BIN_NF64 * sizeof(double), D.stream));
  CK(cudaMemsetAsync(M.counts, 0, nb * R3D_BIN_NCNT * sizeof(unsigned long long), D.stream));
  CK(cudaMemsetAsync(M.counters, 0, R3D_NCOUNTERS * sizeof(unsigned long long), D.stream));

  // launch geometry: persistent grid, a whole number of resident CTAs per SM
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, D.device));
  size_t cell_bytes = nc * d->cell_nparam * sizeof(double);
  D.cells_in_smem = cell_bytes <= 16 * 1024;
  D.smem = (size_t)d->n_seis * sizeof(double4) + (D.cells_in_smem ? cell_bytes : 0);
  if (D.smem > (size_t)prop.sharedMemPerBlockOptin)
    return fail(R3D_EUNSUPPORTED, "too many seismometers for the shared-memory scan table");
  int per_sm = 1;
  for (int trace = 0; trace < 2; trace++) {
    propagate_fn fn = pick_kernel(d->cell_kind, trace != 0);
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)D.smem));
    int occ = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, R3D_THREADS, D.smem));
    if (trace == 0) per_sm = std::max(occ, 1);
  }
  D.grid = prop.multiProcessorCount * per_sm;
  CK(cudaStreamSynchronize(D.stream));
  CK(cudaGetLastError());
  return 0;
}

void destroy_device(DevState &D) {
  if (D.device < 0) return;
  cudaSetDevice(D.device);
  if (D.stream) cudaStreamSynchronize(D.stream);
  for (auto &ev : D.timing) { cudaEventDestroy(ev.first); cudaEventDestroy(ev.second); }
  for (void *p : D.allocs) cudaFree(p);
  if (D.stream) cudaStreamDestroy(D.stream);
  D.allocs.clear(); D.timing.clear(); D.stream = nullptr;
}

int launch(r3d_handle *h, DevState &D, unsigned long long first, unsigned long long n, unsigned long long seed, r3d_phonon_final *finals) {
  CK(cudaSetDevice(D.device));
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  D.timing.push_back({a, b});
  CK(cudaMemsetAsync(D.M.next_phonon, 0, sizeof(unsigned long long), D.stream));
  CK(cudaEventRecord(a, D.stream));
  if (n) {
    // no more CTAs than there is work for (one warp drains R3D_CHUNK phonons at a time)
    unsigned long long want = (n + R3D_THREADS - 1) / R3D_THREADS;
    int grid = (int)std::min<unsigned long long>((unsigned long long)D.grid, std::max<unsigned long long>(want, 1));
    pick_kernel(h->cell_kind, finals != nullptr)<<<grid, R3D_THREADS, D.smem, D.stream>>>(D.M, first, n, seed, D.cells_in_smem, finals);
    h->launches++;
  }
  CK(cudaEventRecord(b, D.stream));
  CK(cudaGetLastError());
  return 0;
}

}  // namespace

// ===========================================================================
// C ABI
// ===========================================================================
extern "C" {

const char *r3d_last_error(void) { return g_err.c_str(); }
int r3d_abi_version(void) { return R3D_ABI_VERSION; }

int r3d_create(const r3d_model_desc *desc, const int *devices, int n_dev, r3d_handle **out) {
  if (!out) return fail(R3D_EINVAL, "null output handle");
  *out = nullptr;
  if (int rc = validate(desc)) return rc;
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail(R3D_ENODEV, std::string("no usable CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  if (n_dev <= 0) return fail(R3D_EINVAL, "n_dev must be >= 1");
  int guide_bits = -1;
  if (const char *s = getenv("R3D_GUIDE_BITS")) guide_bits = atoi(s);
  r3d_handle *h = new r3d_handle();
  h->cell_kind = desc->cell_kind; h->n_seis = desc->n_seis; h->n_bins = desc->n_bins;
  h->devs.resize(n_dev);
  for (int i = 0; i < n_dev; i++) {
    int dev = devices ? devices[i] : i;
    if (dev < 0 || dev >= count) { r3d_destroy(h); return fail(R3D_ENODEV, "device index out of range"); }
    h->devs[i].device = dev;
    if (int rc = build_device(h->devs[i], desc, guide_bits)) { std::string keep = g_err; r3d_destroy(h); g_err = keep; return rc; }
  }
  *out = h;
  return 0;
}

void r3d_destroy(r3d_handle *h) {
  if (!h) return;
  for (auto &D : h->devs) destroy_device(D);
  delete h;
}

int r3d_run(r3d_handle *h, uint64_t first_phonon, uint64_t n_phonons, uint64_t seed) {
  if (!h) return fail(R3D_EINVAL, "null handle");
  const uint64_t G = h->devs.size();
  for (uint64_t g = 0; g < G; g++) {      // contiguous index ranges
    uint64_t lo = n_phonons / G * g + (n_phonons % G) * g / G;
    uint64_t hi = n_phonons / G * (g + 1) + (n_phonons % G) * (g + 1) / G;
    if (int rc = launch(h, h->devs[g], first_phonon + lo, hi - lo, seed, nullptr)) return rc;
  }
  return 0;
}

int r3d_sync(r3d_handle *h, double *device_seconds) {
  if (!h) return fail(R3D_EINVAL, "null handle");
  double worst = 0;
  for (auto &D : h->devs) {
    CK(cudaSetDevice(D.device));
    CK(cudaStreamSynchronize(D.stream));
    double sum = 0;
    for (auto &ev : D.timing) {
      float ms = 0;
      CK(cudaEventElapsedTime(&ms, ev.first, ev.second));
      sum += ms * 1e-3;
      cudaEventDestroy(ev.first); cudaEventDestroy(ev.second);
    }
    D.timing.clear();
    worst = std::max(worst, sum);
  }
  if (device_seconds) *device_seconds = worst;
  return 0;
}

int r3d_fetch(r3d_handle *h, double *energies, uint64_t *counts, uint64_t *counters, uint32_t *diag) {
  if (!h) return fail(R3D_EINVAL, "null handle");
  const size_t nb = (size_t)h->n_seis * h->n_bins;
  std::vector<double> e;
  std::vector<unsigned long long> c;
  unsigned long long k[R3D_NCOUNTERS], ksum[R3D_NCOUNTERS] = {0};
  if (energies) memset(energies, 0, nb * R3D_BIN_NF64 * sizeof(double));
  if (counts) memset(counts, 0, nb * R3D_BIN_NCNT * sizeof(uint64_t));
  for (size_t g = 0; g < h->devs.size(); g++) {
    DevState &D = h->devs[g];
    CK(cudaSetDevice(D.device));
    CK(cudaStreamSynchronize(D.stream));
    if (energies && nb) {
      if (g == 0) CK(cudaMemcpy(energies, D.M.energies, nb * R3D_BIN_NF64 * sizeof(double), cudaMemcpyDeviceToHost));
      else {
        e.resize(nb * R3D_BIN_NF64);
        CK(cudaMemcpy(e.data(), D.M.energies, e.size() * sizeof(double), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < e.size(); i++) energies[i] += e[i];
      }
    }
    if (counts && nb) {
      if (g == 0) CK(cudaMemcpy(counts, D.M.counts, nb * R3D_BIN_NCNT * sizeof(uint64_t), cudaMemcpyDeviceToHost));
      else {
        c.resize(nb * R3D_BIN_NCNT);
        CK(cudaMemcpy(c.data(), D.M.counts, c.size() * sizeof(uint64_t), cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < c.size(); i++) counts[i] += c[i];
      }
    }
    CK(cudaMemcpy(k, D.M.counters, sizeof k, cudaMemcpyDeviceToHost));
    for (int i = 0; i < R3D_NCOUNTERS; i++) { if (i == 7) ksum[i] |= k[i]; else ksum[i] += k[i]; }
  }
  if (counters) for (int i = 0; i < R3D_NCOUNTERS; i++) counters[i] = ksum[i];
  if (diag) *diag = (uint32_t)ksum[7];
  return 0;
}

int r3d_reset(r3d_handle *h) {
  if (!h) return fail(R3D_EINVAL, "null handle");
  const size_t nb = std::max<size_t>((size_t)h->n_seis * h->n_bins, 1);
  for (auto &D : h->devs) {
    CK(cudaSetDevice(D.device));
    CK(cudaMemsetAsync(D.M.energies, 0, nb * R3D_BIN_NF64 * sizeof(double), D.stream));
    CK(cudaMemsetAsync(D.M.counts, 0, nb * R3D_BIN_NCNT * sizeof(unsigned long long), D.stream));
    CK(cudaMemsetAsync(D.M.counters, 0, R3D_NCOUNTERS * sizeof(unsigned long long), D.stream));
  }
  return 0;
}

int r3d_device_accumulators(r3d_handle *h, int dev_slot, void **energies, void **counts, void **counters) {
  if (!h || dev_slot < 0 || dev_slot >= (int)h->devs.size()) return fail(R3D_EINVAL, "bad handle or device slot");
  if (energies) *energies = h->devs[dev_slot].M.energies;
  if (counts) *counts = h->devs[dev_slot].M.counts;
  if (counters) *counters = h->devs[dev_slot].M.counters;
  return 0;
}

int r3d_stream(r3d_handle *h, int dev_slot, void **stream) {
  if (!h || dev_slot < 0 || dev_slot >= (int)h->devs.size() || !stream) return fail(R3D_EINVAL, "bad handle or device slot");
  *stream = h->devs[dev_slot].stream;
  return 0;
}

int r3d_launch_count(r3d_handle *h, uint64_t *n) {
  if (!h || !n) return fail(R3D_EINVAL, "null argument");
  *n = h->launches;
  return 0;
}

int r3d_trace(r3d_handle *h, uint64_t first_phonon, uint64_t n_phonons, uint64_t seed, r3d_phonon_final *out) {
  if (!h || !out) return fail(R3D_EINVAL, "null argument");
  if (!n_phonons) return 0;
  DevState &D = h->devs[0];
  CK(cudaSetDevice(D.device));
  r3d_phonon_final *dfin = nullptr;
  CK(cudaMalloc(&dfin, n_phonons * sizeof(r3d_phonon_final)));
  int rc = launch(h, D, first_phonon, n_phonons, seed, dfin);
  cudaError_t e = cudaSuccess;
  if (!rc) e = cudaStreamSynchronize(D.stream);
  if (!rc && e == cudaSuccess) e = cudaMemcpy(out, dfin, n_phonons * sizeof(r3d_phonon_final), cudaMemcpyDeviceToHost);
  cudaFree(dfin);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(R3D_ECUDA, std::string("r3d_trace: ") + cudaGetErrorString(e));
  return 0;
}

}  // extern "C"

// ---- sub-kernel hooks --------------------------------------------------------
namespace {
struct Scratch {               // device buffers of one hook call, freed on scope exit
  std::vector<void *> p;
  ~Scratch() { for (void *q : p) cudaFree(q); }
  template <class T> int up(const T *host, size_t n, T **dev) {
    void *q = nullptr;
    CK(cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)));
    p.push_back(q);
    if (host && n) CK(cudaMemcpy(q, host, n * sizeof(T), cudaMemcpyHostToDevice));
    *dev = static_cast<T *>(q);
    return 0;
  }
};
int need_device() {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) return fail(R3D_ENODEV, "no usable CUDA device");
  return 0;
}
int path_hook(r3d_handle *h, const double *in, uint32_t n, double *out, int advance_mode) {
  if (!h || !in || !out) return fail(R3D_EINVAL, "null argument");
  DevState &D = h->devs[0];
  CK(cudaSetDevice(D.device));
  Scratch S;
  double *din, *dout;
  if (int rc = S.up(in, (size_t)n * (advance_mode ? 8 : 7), &din)) return rc;
  if (int rc = S.up((const double *)nullptr, (size_t)n * 9, &dout)) return rc;
  unsigned g = (n + 127) / 128;
  switch (h->cell_kind) {
    case R3D_CELL_CYLINDER: test_path_kernel<Cylinder><<<g, 128>>>(D.M, din, n, dout, advance_mode); break;
    case R3D_CELL_SHELL: test_path_kernel<Shell><<<g, 128>>>(D.M, din, n, dout, advance_mode); break;
    default: test_path_kernel<Tetra><<<g, 128>>>(D.M, din, n, dout, advance_mode); break;
  }
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, dout, (size_t)n * 9 * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}
template <class K>
int rows_hook(K launch_fn, const double *in, uint32_t n, int win, int wout, double *out) {
  if (!in || !out) return fail(R3D_EINVAL, "null argument");
  if (int rc = need_device()) return rc;
  Scratch S;
  double *din, *dout;
  if (int rc = S.up(in, (size_t)n * win, &din)) return rc;
  if (int rc = S.up((const double *)nullptr, (size_t)n * wout, &dout)) return rc;
  launch_fn(din, dout);
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, dout, (size_t)n * wout * sizeof(double), cudaMemcpyDeviceToHost));
  return 0;
}
}  // namespace

extern "C" {

int r3d_test_cdf_search(const double *cdf, uint32_t n_cdf, const uint32_t *k, uint32_t n, uint32_t *out, int use_guide_table) {
  if (!cdf || !k || !out || !n_cdf) return fail(R3D_EINVAL, "null argument");
  if (int rc = need_device()) return rc;
  Scratch S;
  double *dc; uint32_t *dk, *dout, *dg;
  if (int rc = S.up(cdf, n_cdf, &dc)) return rc;
  if (int rc = S.up(k, n, &dk)) return rc;
  if (int rc = S.up((const uint32_t *)nullptr, n, &dout)) return rc;
  uint32_t shift = 32, stride = 1;
  if (use_guide_table && n_cdf >= 16) {
    int bits = use_guide_table > 1 ? use_guide_table : 0;
    if (!bits) while ((1ull << (bits + 2)) < n_cdf && bits < 24) bits++;
    if (bits > 24) bits = 24;
    shift = 31 - bits; stride = (1u << bits) + 1;
  }
  if (int rc = S.up((const uint32_t *)nullptr, stride, &dg)) return rc;
  if (shift < 32) build_guide_kernel<<<(stride + 255) / 256, 256>>>(dc, n_cdf, 1, shift, stride, dg);
  test_cdf_kernel<<<(n + 127) / 128, 128>>>(dc, n_cdf, dg, shift, dk, n, dout);
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, dout, n * sizeof(uint32_t), cudaMemcpyDeviceToHost));
  return 0;
}

int r3d_test_path_to_boundary(r3d_handle *h, const double *in, uint32_t n, double *out) { return path_hook(h, in, n, out, 0); }
int r3d_test_advance(r3d_handle *h, const double *in, uint32_t n, double *out) { return path_hook(h, in, n, out, 1); }

int r3d_test_transform(const double *in, uint32_t n, double *out) {
  return rows_hook([&](double *a, double *b) { test_transform_kernel<<<(n + 127) / 128, 128>>>(a, n, b); }, in, n, 6, 3, out);
}
int r3d_test_rtcoef(const double *in, uint32_t n, double *out) {
  return rows_hook([&](double *a, double *b) { test_rtcoef_kernel<<<(n + 127) / 128, 128>>>(a, n, b); }, in, n, 15, 13, out);
}
int r3d_test_catch(double bin_dt, uint32_t n_bins, const double *in, uint32_t n, double *out) {
  return rows_hook([&](double *a, double *b) { test_catch_kernel<<<(n + 127) / 128, 128>>>(bin_dt, n_bins, a, n, b); }, in, n, 28, 6, out);
}

}  // extern "C"
