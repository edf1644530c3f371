// r3d_wavefront.cuh -- the propagate loop as a wavefront over a pool of phonons in HBM.
//
// Why: one iteration of the reference's Propagate loop (phonons.cpp:542) is one of a few very different events
// (scatter / cell-to-cell hand-over / surface reflection with a seismometer scan / loss / time-out / new phonon).
// With one phonon per thread in one fused kernel a warp holds all of them at once; ncu measured 5.3 of 32 lanes
// active per instruction and an instruction-fetch stall of 8.8 cycles per issue (profiles/r1_fused_kernel.md).
// Here the phonon state lives in a struct-of-arrays pool in HBM and every step runs three small kernels, with the
// phonons regrouped in between by index queues, so that each warp executes ONE kind of event:
//   advance   (one thread per pool slot)    refill dead slots from the work counter; time-out / validity checks;
//                                           distance to boundary; path-length draw; move; cheap hand-overs inline
//   draw      (one thread per queued draw)  exact guide-table CDF search, take-off-angle fetch, then either
//                                           source-phonon initialisation or Phonon::Transform
//   interface (one thread per queued hit)   seismometer catch through a uniform-grid index, then R/T coefficients /
//                                           ray bending
// The price is HBM traffic for the state (~300 B per event), which is what a B200 has to spare on this workload.
#pragma once
#include "r3d_device.cuh"

namespace r3d {

#define R3D_FULL 0xffffffffu

// ---- the pool (struct of arrays; one slot = one phonon in flight, phonons.hpp:69-126) -----------------------
struct Pool {
  uint32_t n_slots;
  double *time, *pathlen, *recent, *amp, *lx, *ly, *lz, *th, *ph, *pol;
  uint32_t *moves, *cell, *ordinal;
  uint8_t *type, *alive;
  unsigned long long *idx;
  uint32_t *tr_catches, *tr_scatters, *tr_iters;   // per-phonon statistics, trace mode only
  // queues (rebuilt every step)
  uint4 *q_draw;        // {slot, 31-bit draw, table index, kind}: kind 0 = source take-off angle, 1 = scatter angle
  uint2 *q_face;        // {slot, exit face}
  uint32_t *q_count;    // [0] draws, [1] faces, [2] "some slot is alive" flag
  unsigned long long *block_tally;   // [n_blocks_max][R3D_NCOUNTERS], owned per block: no atomics
};

struct Job { unsigned long long first, n, seed; r3d_phonon_final *finals; };

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

// Rng positioned at an arbitrary ordinal (state is only the ordinal; the block is recomputed when needed)
struct RngAt {
  Rng g; bool have;
  R3D_DEV void init(unsigned long long seed, unsigned long long idx, uint32_t ordinal) { g.init(seed, idx); g.ordinal = ordinal; have = false; }
  R3D_DEV uint32_t next() {
    if (!have || (g.ordinal & 3u) == 0u) { g.block(g.ordinal >> 2); have = true; }
    uint32_t o = g.ordinal & 3u;
    uint32_t v = (o == 0) ? g.w[0] : (o == 1) ? g.w[1] : (o == 2) ? g.w[2] : g.w[3];
    g.ordinal++;
    return v >> 1;
  }
};

// per-thread tallies -> per-block row (dataout.cpp:591-617).  Every block owns one row of block_tally, so the
// read-modify-write needs no atomics; rows are summed at fetch time.
struct Tally {
  unsigned long long v[R3D_NCOUNTERS];
  R3D_DEV void clear() {
#pragma unroll
    for (int i = 0; i < R3D_NCOUNTERS; i++) v[i] = 0;
  }
  R3D_DEV void died(uint32_t fate) {
    switch (fate & 0xFF) {
      case R3D_FATE_LOST: v[R3D_CNT_LOST]++; break;
      case R3D_FATE_TIMEOUT: v[R3D_CNT_TIMEOUT]++; break;
      default: v[R3D_CNT_INVALID]++; v[7] |= (fate >> 8); break;
    }
  }
  // all threads of the block must call this
  R3D_DEV void flush(unsigned long long *row, unsigned long long (*sm)[R3D_NCOUNTERS]) {
    const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int k = 0; k < R3D_NCOUNTERS; k++) {
      unsigned long long x = v[k];
      if (k == 7) { for (int o = 16; o > 0; o >>= 1) x |= __shfl_down_sync(R3D_FULL, x, o); }
      else { for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(R3D_FULL, x, o); }
      if (lane == 0) sm[warp][k] = x;
    }
    __syncthreads();
    if (threadIdx.x < R3D_NCOUNTERS) {
      unsigned long long x = 0;
      for (unsigned w = 0; w < nw; w++) { if (threadIdx.x == 7) x |= sm[w][threadIdx.x]; else x += sm[w][threadIdx.x]; }
      if (x) { if (threadIdx.x == 7) row[threadIdx.x] |= x; else row[threadIdx.x] += x; }
    }
  }
};

template <bool TRACE>
R3D_DEV void write_final(const Pool &Q, const Job &J, uint32_t s, const Phonon &p, uint32_t fate, uint32_t ordinal) {
  if (!TRACE) return;
  r3d_phonon_final *f = J.finals + (Q.idx[s] - J.first);
  f->time = p.time; f->pathlen = p.pathlen; f->amp = p.amp;
  f->loc[0] = p.loc.x; f->loc[1] = p.loc.y; f->loc[2] = p.loc.z;
  f->theta = p.th; f->phi = p.ph; f->pol = p.pol;
  f->moves = p.moves; f->cell = p.cell; f->type = (uint32_t)p.type; f->fate = fate;
  f->draws = ordinal; f->catches = Q.tr_catches[s]; f->scatters = Q.tr_scatters[s]; f->iters = Q.tr_iters[s];
}

// warp-aggregated queue append: one atomic per warp, consecutive positions for the lanes that push
R3D_DEV uint32_t queue_slot(uint32_t *counter, bool push) {
  const unsigned mask = __ballot_sync(R3D_FULL, push);
  if (!mask) return 0;
  const unsigned lane = threadIdx.x & 31u;
  const int leader = __ffs(mask) - 1;
  uint32_t base = 0;
  if ((int)lane == leader) base = atomicAdd(counter, (uint32_t)__popc(mask));
  base = __shfl_sync(R3D_FULL, base, leader);
  return base + __popc(mask & ((1u << lane) - 1u));
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

// =====================================================================================================
// kernel A: advance.  Grid-stride over tiles of blockDim.x pool slots.
// =====================================================================================================
#define R3D_A_THREADS 256
template <class Cell, bool TRACE>
__global__ void __launch_bounds__(R3D_A_THREADS)
advance_kernel(const DevModel M, const Pool Q, const Job J, int cells_in_smem) {
  extern __shared__ double smem_cells[];
  __shared__ unsigned long long tally_sm[R3D_A_THREADS / 32][R3D_NCOUNTERS];
  __shared__ uint32_t warp_dead[R3D_A_THREADS / 32];
  __shared__ unsigned long long tile_base;
  if (cells_in_smem)
    for (uint32_t i = threadIdx.x; i < M.n_cells * M.cell_nparam; i += blockDim.x) smem_cells[i] = M.cell_params[i];
  __syncthreads();
  const double *cells = cells_in_smem ? smem_cells : M.cell_params;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  Tally T; T.clear();
  bool any_alive = false;

  const uint32_t n_tiles = (Q.n_slots + blockDim.x - 1) / blockDim.x;
  for (uint32_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const uint32_t s = tile * blockDim.x + threadIdx.x;
    const bool in_range = s < Q.n_slots;
    bool alive = in_range && Q.alive[s];

    // ---- refill dead slots: one atomic on the work counter per tile -----------------------------------
    const bool dead = in_range && !alive;
    const unsigned dmask = __ballot_sync(R3D_FULL, dead);
    if (lane == 0) warp_dead[warp] = __popc(dmask);
    __syncthreads();
    if (threadIdx.x == 0) {
      uint32_t tot = 0;
      for (unsigned w = 0; w < blockDim.x / 32; w++) { uint32_t c = warp_dead[w]; warp_dead[w] = tot; tot += c; }
      unsigned long long b = J.n;                       // nothing to hand out unless the counter says so
      if (tot && *(volatile unsigned long long *)M.next_phonon < J.n) b = atomicAdd(M.next_phonon, (unsigned long long)tot);
      tile_base = b;
    }
    __syncthreads();
    uint32_t fate = 0;
    bool push_draw = false, push_face = false;
    uint32_t draw_k = 0, draw_table = 0, draw_kind = 1u, face = 0;
    if (dead) {
      const unsigned long long cand = tile_base + warp_dead[warp] + __popc(dmask & ((1u << lane) - 1u));
      if (cand < J.n) {
        // ShearDislocation::GenerateEventPhonon (events.cpp:111-124) -> Phonon ctor (phonons.hpp:193-207)
        const unsigned long long idx = J.first + cand;
        RngAt g; g.init(J.seed, idx, 0);
        const uint32_t rt3 = cdf_search_small(M.src_whole, 3, g.next());
        draw_k = g.next();
        draw_table = rt3; draw_kind = 0u; push_draw = true;     // the take-off angle is drawn by this step's draw kernel
        Q.time[s] = 0; Q.pathlen[s] = 0; Q.recent[s] = 0; Q.amp[s] = 1.0;
        Q.lx[s] = M.src_loc[0]; Q.ly[s] = M.src_loc[1]; Q.lz[s] = M.src_loc[2];
        Q.pol[s] = (rt3 == R3D_RAY_SH) ? kPi * 0.5 : 0.0;
        Q.type[s] = (rt3 == R3D_RAY_P) ? R3D_RAY_P : R3D_RAY_S;
        Q.moves[s] = 0; Q.cell[s] = M.src_cell; Q.ordinal[s] = 2; Q.idx[s] = idx; Q.alive[s] = 1;
        if (TRACE) { Q.tr_catches[s] = 0; Q.tr_scatters[s] = 0; Q.tr_iters[s] = 0; }
        T.v[6]++;
        any_alive = true;
      }
    }

    // ---- one Propagate-loop iteration up to the event's classification (phonons.cpp:542-623) -------------
    Phonon p;
    uint32_t ordinal = 0;
    if (alive) {
      p.time = Q.time[s]; p.pathlen = Q.pathlen[s]; p.recent = Q.recent[s]; p.amp = Q.amp[s];
      p.loc = V(Q.lx[s], Q.ly[s], Q.lz[s]);
      p.th = Q.th[s]; p.ph = Q.ph[s]; p.pol = Q.pol[s];
      p.moves = Q.moves[s]; p.cell = Q.cell[s]; p.type = Q.type[s];
      ordinal = Q.ordinal[s];
      T.v[R3D_CNT_EVENTS]++;
      if (TRACE) Q.tr_iters[s]++;
      bool recent_reset = false;
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
        else { p.recent = 0; recent_reset = true; }
      }
      (void)recent_reset;
      if (!fate) {
        const double *c = cells + (size_t)p.cell * M.cell_nparam;
        typename Cell::Path P;
        const double edgelen = Cell::path(M, c, p.type, p.loc, p.th, p.ph, P);
        if (edgelen == pinf()) fate = R3D_FATE_TIMEOUT;       // phonons.cpp:595-598
        else {
          RngAt g; g.init(J.seed, Q.idx[s], ordinal);
          const uint32_t scat = __ldg(M.cell_scat + p.cell);
          // Scatterer::GetRandomPathLength (scatterers.cpp:297-307)
          const double r = 1.0 - ((double)g.next()) / (kRandMax + 1);
          const double scatlen = -log(r) * __ldg(M.scat_mfp + scat * 2 + p.type);
          const bool scatter = scatlen < edgelen;
          Travel tr = Cell::advance(M, c, p.type, scatter ? scatlen : edgelen, p.loc, p.th, p.ph, P);
          move(p, tr);
          if (scatter) {
            // Scatterer::GetRandomScatteredRelativePhonon (scatterers.cpp:318-363); the table draw is queued
            if (M.no_deflect) {
              transform(p.th, p.ph, p.pol, M.min_theta, 0.0, 0.0);
              T.v[R3D_CNT_SCATTERS]++;
              if (TRACE) Q.tr_scatters[s]++;
            } else {
              const uint32_t conv = cdf_search_small(M.scat_whole + (scat * 2 + p.type) * 4, 4, g.next());
              draw_k = g.next();
              draw_table = scat * 4 + conv;
              push_draw = true;
            }
          } else {
            const uint32_t fi = p.cell * M.faces_per_cell + P.face;
            const uint32_t fl = __ldg(M.face_flags + fi);
            face = P.face;
            if (fl & (R3D_FACE_COLLECT | R3D_FACE_REFLECT)) push_face = true;
            else if (fl & R3D_FACE_ADJOIN) {                   // Phonon::Refract (phonons.cpp:225-255)
              const uint32_t other = __ldg(M.face_other + fi);
              if ((fl & R3D_FACE_DISCON) || velocity_jump<Cell>(M, cells, p.cell, other, p.loc) > 0.00001) push_face = true;
              else p.cell = other;                             // Refraction_Continuous
            } else fate = R3D_FATE_LOST;                       // phonons.cpp:675
          }
          ordinal = g.g.ordinal;
          Q.time[s] = p.time; Q.pathlen[s] = p.pathlen; Q.amp[s] = p.amp;
          Q.lx[s] = p.loc.x; Q.ly[s] = p.loc.y; Q.lz[s] = p.loc.z;
          Q.th[s] = p.th; Q.ph[s] = p.ph; Q.pol[s] = p.pol;
          Q.moves[s] = p.moves; Q.cell[s] = p.cell; Q.ordinal[s] = ordinal;
        }
      }
      Q.recent[s] = p.recent;
      if (fate) {
        Q.alive[s] = 0;
        T.died(fate);
        write_final<TRACE>(Q, J, s, p, fate, ordinal);
      } else any_alive = true;
    }
    // queue appends (warp-uniform calls)
    uint32_t at = queue_slot(Q.q_count + 0, push_draw);
    if (push_draw) Q.q_draw[at] = make_uint4(s, draw_k, draw_table, draw_kind);
    at = queue_slot(Q.q_count + 1, push_face);
    if (push_face) Q.q_face[at] = make_uint2(s, face);
  }
  if (__syncthreads_or(any_alive) && threadIdx.x == 0) Q.q_count[2] = 1;
  T.flush(Q.block_tally + (size_t)blockIdx.x * R3D_NCOUNTERS, tally_sm);
}

// =====================================================================================================
// kernel B: draw.  ProbDist::GetRandomIndex on the queued table + take-off angle, then either the new
// phonon's direction (sources.cpp:156-170) or Phonon::Transform (phonons.cpp:116-170).
// =====================================================================================================
#define R3D_B_THREADS 256
template <bool TRACE>
__global__ void __launch_bounds__(R3D_B_THREADS)
draw_kernel(const DevModel M, const Pool Q, uint32_t tally_row0) {
  __shared__ unsigned long long tally_sm[R3D_B_THREADS / 32][R3D_NCOUNTERS];
  Tally T; T.clear();
  const uint32_t n = Q.q_count[0];
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint4 q = Q.q_draw[i];
    const uint32_t s = q.x;
    if (q.w == 0u) {
      const uint32_t ti = cdf_search(M.src_cdf + (size_t)q.z * M.n_toa, M.n_toa, M.src_guide + (size_t)q.z * M.guide_stride, M.guide_shift, q.y);
      const double2 t = __ldg(M.toa + ti);
      Q.th[s] = t.x; Q.ph[s] = t.y;
    } else {
      const uint32_t ti = cdf_search(M.scat_cdf + (size_t)q.z * M.n_toa, M.n_toa, M.scat_guide + (size_t)q.z * M.guide_stride, M.guide_shift, q.y);
      const double2 t = __ldg(M.toa + ti);
      const uint32_t conv = q.z & 3u;
      const double rpol = (conv == 3u) ? __ldg(M.scat_spol + (size_t)(q.z >> 2) * M.n_toa + ti) : 0.0;
      double th = Q.th[s], ph = Q.ph[s], pol = Q.pol[s];
      transform(th, ph, pol, t.x, t.y, rpol);
      Q.th[s] = th; Q.ph[s] = ph; Q.pol[s] = pol;
      Q.type[s] = (uint8_t)(conv & 1u);                       // PP,PS,SP,SS -> P,S,P,S
      T.v[R3D_CNT_SCATTERS]++;
      if (TRACE) Q.tr_scatters[s]++;
    }
  }
  T.flush(Q.block_tally + (size_t)(tally_row0 + blockIdx.x) * R3D_NCOUNTERS, tally_sm);
}

// =====================================================================================================
// kernel C: interface.  Everything that happens at a face that is not a plain hand-over:
// collection (dataout.cpp:545-568, 103-216), free-surface / discontinuity R/T (phonons.cpp:429-476),
// Snell bending (phonons.cpp:311-405).
// =====================================================================================================
#define R3D_C_THREADS 128

// Phonon::Refraction_FullRT + CellFace::GetRTBasis (media_cellface.cpp:122-149)
template <class Cell>
R3D_DEV void refraction_fullrt(const DevModel &M, const double *cells, Phonon &p, int face, bool adjoin, uint32_t other, RngAt &g) {
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

template <class Cell, bool TRACE>
__global__ void __launch_bounds__(R3D_C_THREADS, 4)
interface_kernel(const DevModel M, const Pool Q, const Job J, int cells_in_smem, uint32_t tally_row0) {
  extern __shared__ double4 smem_c4[];
  __shared__ unsigned long long tally_sm[R3D_C_THREADS / 32][R3D_NCOUNTERS];
  double4 *sph = smem_c4;                                     // [n_seis] (x, y, z, r_out^2 (1+eps))
  double *scells = reinterpret_cast<double *>(smem_c4 + M.n_seis);
  const uint32_t n = Q.q_count[1];
  Tally T; T.clear();
  if (blockIdx.x * blockDim.x < n) {                          // blocks without work skip the table load
    for (uint32_t i = threadIdx.x; i < M.n_seis; i += blockDim.x) sph[i] = M.seis_sphere[i];
    if (cells_in_smem)
      for (uint32_t i = threadIdx.x; i < M.n_cells * M.cell_nparam; i += blockDim.x) scells[i] = M.cell_params[i];
  }
  __syncthreads();
  const double *cells = cells_in_smem ? scells : M.cell_params;

  const uint32_t n_round = (n + 31u) & ~31u;                  // whole warps stay in the loop together
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_round; i += gridDim.x * blockDim.x) {
    const bool have = i < n;
    uint32_t s = 0, fl = 0, other = 0, ordinal = 0;
    int face = 0;
    Phonon p;
    p.time = p.pathlen = p.recent = p.amp = 0; p.loc = V(0, 0, 0); p.th = p.ph = p.pol = 0; p.moves = 0; p.cell = 0; p.type = 0;
    if (have) {
      const uint2 q = Q.q_face[i];
      s = q.x; face = (int)q.y;
      p.time = Q.time[s]; p.amp = Q.amp[s];
      p.loc = V(Q.lx[s], Q.ly[s], Q.lz[s]);
      p.th = Q.th[s]; p.ph = Q.ph[s]; p.pol = Q.pol[s];
      p.cell = Q.cell[s]; p.type = Q.type[s];
      ordinal = Q.ordinal[s];
      const uint32_t fi = p.cell * M.faces_per_cell + face;
      fl = __ldg(M.face_flags + fi);
      other = __ldg(M.face_other + fi);
    }

    // ---- collection (dataout.cpp:545-568): every seismometer is pass-through (dataout.cpp:50), so all that contain
    // the point must bin it.  Candidates come from the uniform grid over the seismometers' bounding spheres. --------
    uint32_t my_catches = 0;
    if (have && (fl & R3D_FACE_COLLECT) && M.n_seis > 0) {
      const int cx_ = grid_axis_cell(p.loc.x, M.grid_min[0], M.grid_inv_h[0]);
      const int cy_ = grid_axis_cell(p.loc.y, M.grid_min[1], M.grid_inv_h[1]);
      const int cz_ = grid_axis_cell(p.loc.z, M.grid_min[2], M.grid_inv_h[2]);
      if (cx_ >= 0 && cy_ >= 0 && cz_ >= 0 && cx_ < (int)M.grid_dim[0] && cy_ < (int)M.grid_dim[1] && cz_ < (int)M.grid_dim[2]) {
        const uint32_t cell = ((uint32_t)cz_ * M.grid_dim[1] + (uint32_t)cy_) * M.grid_dim[0] + (uint32_t)cx_;
        const uint32_t i0 = __ldg(M.grid_start + cell), i1 = __ldg(M.grid_start + cell + 1);
        for (uint32_t j = i0; j < i1; j++) {
          const uint32_t k2 = __ldg(M.grid_items + j);
          const double4 q = sph[k2];
          const double dx = q.x - p.loc.x, dy = q.y - p.loc.y, dz = q.z - p.loc.z;
          if (dx * dx + dy * dy + dz * dz > q.w) continue;       // cannot be within the gather radius
          // rare from here on (a few per cent of the surface hits): the exact CatchPhonon test
          const double vel = Cell::veloc(cells + (size_t)p.cell * M.cell_nparam, p.type, p.loc);
          const v3 dir = from_thph(p.th, p.ph);
          const v3 dopm = dir_of_motion(p.type, p.th, p.ph, p.pol);
          uint32_t bin; double e[4];
          if (seis_catch(M.seis + (size_t)k2 * R3D_SEIS_NPARAM, M.bin_dt, M.n_bins, p.time, p.loc, dir, dopm, p.type, p.amp, vel, bin, e)) {
            const size_t b = (size_t)k2 * M.n_bins + bin;
            atomicAdd(M.energies + b * 5 + 0, e[0]);
            atomicAdd(M.energies + b * 5 + 1, e[1]);
            atomicAdd(M.energies + b * 5 + 2, e[2]);
            atomicAdd(M.energies + b * 5 + 3 + p.type, e[3]);
            atomicAdd(M.counts + b * 2 + p.type, 1ull);
            my_catches++;
          }
        }
      }
    }
    if (have) {
      T.v[R3D_CNT_CATCHES] += my_catches;
      if (TRACE && my_catches) Q.tr_catches[s] += my_catches;

      // ---- reflection / refraction (phonons.cpp:640-676) ---------------------------------------------------
      uint32_t fate = 0;
      RngAt g; g.init(J.seed, Q.idx[s], ordinal);
      if (fl & R3D_FACE_REFLECT) refraction_fullrt<Cell>(M, cells, p, face, (fl & R3D_FACE_ADJOIN) != 0, other, g);
      else if (fl & R3D_FACE_ADJOIN) {
        if (fl & R3D_FACE_DISCON) refraction_fullrt<Cell>(M, cells, p, face, true, other, g);
        else if (velocity_jump<Cell>(M, cells, p.cell, other, p.loc) > 0.00001) refraction_bend<Cell>(M, cells, p, face, other);
        else p.cell = other;
      } else fate = R3D_FATE_LOST;
      if (fate) {
        Q.alive[s] = 0;
        T.died(fate);
        if (TRACE) { p.pathlen = Q.pathlen[s]; p.moves = Q.moves[s]; }
        write_final<TRACE>(Q, J, s, p, fate, g.g.ordinal);
      } else {
        Q.th[s] = p.th; Q.ph[s] = p.ph; Q.pol[s] = p.pol;
        Q.cell[s] = p.cell; Q.type[s] = (uint8_t)p.type; Q.ordinal[s] = g.g.ordinal;
      }
    }
  }
  T.flush(Q.block_tally + (size_t)(tally_row0 + blockIdx.x) * R3D_NCOUNTERS, tally_sm);
}

}  // namespace r3d
