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
  double *time, *pathlen, *recent, *aexp;      // aexp: amplitude = exp(-aexp)
  double *lx, *ly, *lz;                        // location
  double *dx, *dy, *dz;                        // unit direction of travel            (e3)
  double *sx, *sy, *sz;                        // unit polarisation direction         (s1; carried for P phonons too,
                                               //  exactly as the reference carries mPol, phonons.hpp:109-118)
  uint32_t *moves, *cell, *ordinal;
  uint8_t *type, *alive;
  unsigned long long *idx;
  uint32_t *tr_catches, *tr_scatters, *tr_iters;   // per-phonon statistics, trace mode only
  // queues (rebuilt every step).  Both are filled from the two ends, so that the two kinds of entry never share a
  // warp: scatter draws grow from index 0, source draws from index n_slots-1; P face hits from 0, S from n_slots-1.
  uint4 *q_draw;        // {slot, 31-bit draw, table index, kind}: kind 0 = source take-off angle, 1 = scatter angle
  uint2 *q_face;        // {slot, exit face}
  uint32_t *q_count;    // R3D_Q_* below
  unsigned long long *block_tally;   // [n_blocks_max][R3D_NCOUNTERS], owned per block: no atomics
};

enum { R3D_Q_DRAW_SCAT = 0, R3D_Q_DRAW_SRC, R3D_Q_FACE_P, R3D_Q_FACE_S, R3D_Q_ALIVE, R3D_Q_CURSOR_DRAW, R3D_Q_CURSOR_FACE, R3D_Q_NCOUNT = 8 };

struct Job { unsigned long long first, n, seed; r3d_phonon_final *finals; };

struct Phonon {
  double time, pathlen, recent, aexp;
  v3 loc, dir, s1;
  uint32_t moves, cell;
  int type;
};

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
  // barrier-free variant: each warp adds its sums to the block's row with atomics (rows are per block, so the only
  // contention is between the few warps of one block, once per launch)
  R3D_DEV void flush_warp(unsigned long long *row) {
    const unsigned lane = threadIdx.x & 31u;
#pragma unroll
    for (int k = 0; k < R3D_NCOUNTERS; k++) {
      unsigned long long x = v[k];
      if (k == 7) { for (int o = 16; o > 0; o >>= 1) x |= __shfl_down_sync(R3D_FULL, x, o); }
      else { for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(R3D_FULL, x, o); }
      if (lane == 0 && x) { if (k == 7) atomicOr(row + k, x); else atomicAdd(row + k, x); }
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
  f->time = p.time; f->pathlen = p.pathlen; f->amp = exp(-p.aexp);
  f->loc[0] = p.loc.x; f->loc[1] = p.loc.y; f->loc[2] = p.loc.z;
  angles_of(p.dir, f->theta, f->phi);
  f->pol = pol_angle_of(p.dir, p.s1);
  f->moves = p.moves; f->cell = p.cell; f->type = (uint32_t)p.type; f->fate = fate;
  f->draws = ordinal; f->catches = Q.tr_catches[s]; f->scatters = Q.tr_scatters[s]; f->iters = Q.tr_iters[s];
}

// one 32-byte take-off-angle record through the read-only path (two 16-byte loads of the same sector)
R3D_DEV double4 load_toa(const double4 *p) {
  const double2 a = __ldg(reinterpret_cast<const double2 *>(p)), b = __ldg(reinterpret_cast<const double2 *>(p) + 1);
  return make_double4(a.x, a.y, b.x, b.y);
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
// two-ended queue of capacity cap: kind A entries at [0, nA), kind B entries at (cap - 1 - j), j in [0, nB)
R3D_DEV uint32_t queue_slot2(uint32_t *counterA, uint32_t *counterB, uint32_t cap, bool pushA, bool pushB) {
  const uint32_t a = queue_slot(counterA, pushA);
  const uint32_t b = queue_slot(counterB, pushB);
  return pushA ? a : cap - 1u - b;
}
// A warp pulls the next chunk of 32 logical entries of a two-ended queue.  Chunks [0, cA) cover kind A, chunks
// [cA, cA + cB) kind B, so no warp mixes kinds.  Returns false when the queue is drained; else `at` is this lane's
// entry index (valid iff `have`) and kindB tells which end it came from.
R3D_DEV bool pull_chunk(uint32_t *cursor, uint32_t nA, uint32_t nB, uint32_t cap, uint32_t &at, bool &have, bool &kindB) {
  const unsigned lane = threadIdx.x & 31u;
  const uint32_t cA = (nA + 31u) >> 5, cB = (nB + 31u) >> 5;
  uint32_t c = 0;
  if (lane == 0) c = atomicAdd(cursor, 1u);
  c = __shfl_sync(R3D_FULL, c, 0);
  if (c >= cA + cB) return false;
  kindB = c >= cA;
  const uint32_t j = (kindB ? c - cA : c) * 32u + lane;
  have = j < (kindB ? nB : nA);
  at = kindB ? cap - 1u - j : j;
  return true;
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
#ifndef R3D_A_MINBLOCKS
#define R3D_A_MINBLOCKS 4      // 64 registers: the kernel waits on HBM loads of the pool, resident warps hide them
#endif
template <class Cell, bool TRACE>
__global__ void __launch_bounds__(R3D_A_THREADS, R3D_A_MINBLOCKS)
advance_kernel(const DevModel M, const Pool Q, const Job J, int cells_in_smem) {
  extern __shared__ double smem_cells[];
  __shared__ unsigned long long tally_sm[R3D_A_THREADS / 32][R3D_NCOUNTERS];
  __shared__ uint32_t warp_dead[R3D_A_THREADS / 32];
  __shared__ uint32_t warp_push[R3D_A_THREADS / 32][4], push_base[4];
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
    int face_type = 0;
    if (dead) {
      const unsigned long long cand = tile_base + warp_dead[warp] + __popc(dmask & ((1u << lane) - 1u));
      if (cand < J.n) {
        // ShearDislocation::GenerateEventPhonon (events.cpp:111-124) -> Phonon ctor (phonons.hpp:193-207)
        const unsigned long long idx = J.first + cand;
        RngAt g; g.init(J.seed, idx, 0);
        const uint32_t rt3 = cdf_search_small(M.src_whole, 3, g.next());
        draw_k = g.next();
        draw_table = rt3; draw_kind = 0u; push_draw = true;     // the take-off angle is drawn by this step's draw kernel
        Q.time[s] = 0; Q.pathlen[s] = 0; Q.recent[s] = 0; Q.aexp[s] = 0.0;
        Q.lx[s] = M.src_loc[0]; Q.ly[s] = M.src_loc[1]; Q.lz[s] = M.src_loc[2];
        Q.type[s] = (rt3 == R3D_RAY_P) ? R3D_RAY_P : R3D_RAY_S;    // direction + polarisation: this step's draw kernel
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
      p.time = Q.time[s]; p.pathlen = Q.pathlen[s]; p.recent = Q.recent[s]; p.aexp = Q.aexp[s];
      p.loc = V(Q.lx[s], Q.ly[s], Q.lz[s]);
      p.dir = V(Q.dx[s], Q.dy[s], Q.dz[s]);
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
      if (!fate) {
        const double *c = cells + (size_t)p.cell * M.cell_nparam;
        typename Cell::Path P;
        const double edgelen = Cell::path(M, c, p.type, p.loc, p.dir, P);
        if (edgelen == pinf()) fate = R3D_FATE_TIMEOUT;       // phonons.cpp:595-598
        else {
          RngAt g; g.init(J.seed, Q.idx[s], ordinal);
          const uint32_t scat = __ldg(M.cell_scat + p.cell);
          // Scatterer::GetRandomPathLength (scatterers.cpp:297-307)
          const double r = 1.0 - ((double)g.next()) / (kRandMax + 1);
          const double scatlen = -log(r) * __ldg(M.scat_mfp + scat * 2 + p.type);
          const bool scatter = scatlen < edgelen;
          const Travel tr = Cell::advance(M, c, p.type, scatter ? scatlen : edgelen, p.loc, p.dir, P);
          // Phonon::Move (phonons.cpp:62-70)
          p.pathlen += tr.len; p.time += tr.time; p.recent += tr.time;
          p.loc = tr.loc; p.aexp += tr.aexp; p.moves += 1;
          if (Cell::curved) {            // the ray turned: the polarisation ANGLE is what the reference carries along
            const v3 s1 = V(Q.sx[s], Q.sy[s], Q.sz[s]);
            const v3 ns1 = carry_pol(p.dir, s1, tr.dir);
            Q.sx[s] = ns1.x; Q.sy[s] = ns1.y; Q.sz[s] = ns1.z;
            p.s1 = ns1;
            p.dir = tr.dir;
            Q.dx[s] = p.dir.x; Q.dy[s] = p.dir.y; Q.dz[s] = p.dir.z;
          }
          if (scatter) {
            // Scatterer::GetRandomScatteredRelativePhonon (scatterers.cpp:318-363); the table draw is queued
            if (M.no_deflect) {
              if (!Cell::curved) p.s1 = V(Q.sx[s], Q.sy[s], Q.sz[s]);
              double st, ct;
              sincos(M.min_theta, &st, &ct);                  // Phonon(ThetaPhi(0,0)) nudged to min_theta, pol 0
              transform(p.dir, p.s1, st, ct, 0.0, 1.0, 0.0, 1.0);
              Q.dx[s] = p.dir.x; Q.dy[s] = p.dir.y; Q.dz[s] = p.dir.z;
              Q.sx[s] = p.s1.x; Q.sy[s] = p.s1.y; Q.sz[s] = p.s1.z;
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
            face_type = p.type;
            if (fl & (R3D_FACE_COLLECT | R3D_FACE_REFLECT)) push_face = true;
            else if (fl & R3D_FACE_ADJOIN) {                   // Phonon::Refract (phonons.cpp:225-255)
              const uint32_t other = __ldg(M.face_other + fi);
              if ((fl & R3D_FACE_DISCON) || velocity_jump<Cell>(M, cells, p.cell, other, p.loc) > 0.00001) push_face = true;
              else p.cell = other;                             // Refraction_Continuous
            } else fate = R3D_FATE_LOST;                       // phonons.cpp:675
          }
          ordinal = g.g.ordinal;
          Q.time[s] = p.time; Q.pathlen[s] = p.pathlen; Q.aexp[s] = p.aexp;
          Q.lx[s] = p.loc.x; Q.ly[s] = p.loc.y; Q.lz[s] = p.loc.z;
          Q.moves[s] = p.moves; Q.cell[s] = p.cell; Q.ordinal[s] = ordinal;
        }
      }
      if (recent_reset || !fate) Q.recent[s] = p.recent;
      if (fate) {
        Q.alive[s] = 0;
        T.died(fate);
        if (TRACE) p.s1 = V(Q.sx[s], Q.sy[s], Q.sz[s]);
        write_final<TRACE>(Q, J, s, p, fate, ordinal);
      } else any_alive = true;
    }
    // queue appends, aggregated over the tile: one atomic per queue end per 256 slots
    {
      const int qk = push_draw ? (draw_kind == 1u ? R3D_Q_DRAW_SCAT : R3D_Q_DRAW_SRC)
                   : push_face ? (face_type == R3D_RAY_P ? R3D_Q_FACE_P : R3D_Q_FACE_S) : -1;
      unsigned mine = 0;
#pragma unroll
      for (int k = 0; k < 4; k++) {
        const unsigned m = __ballot_sync(R3D_FULL, qk == k);
        if (lane == 0) warp_push[warp][k] = __popc(m);
        if (qk == k) mine = m;
      }
      __syncthreads();
      if (threadIdx.x < 4) {
        uint32_t tot = 0;
        for (unsigned w = 0; w < blockDim.x / 32; w++) { const uint32_t c = warp_push[w][threadIdx.x]; warp_push[w][threadIdx.x] = tot; tot += c; }
        push_base[threadIdx.x] = tot ? atomicAdd(Q.q_count + threadIdx.x, tot) : 0u;
      }
      __syncthreads();
      if (qk >= 0) {
        const uint32_t j = push_base[qk] + warp_push[warp][qk] + __popc(mine & ((1u << lane) - 1u));
        const uint32_t at = (qk == R3D_Q_DRAW_SCAT || qk == R3D_Q_FACE_P) ? j : Q.n_slots - 1u - j;
        if (push_draw) Q.q_draw[at] = make_uint4(s, draw_k, draw_table, draw_kind);
        else Q.q_face[at] = make_uint2(s, face);
      }
    }
  }
  if (__syncthreads_or(any_alive) && threadIdx.x == 0) Q.q_count[R3D_Q_ALIVE] = 1;
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
  Tally T; T.clear();
  const uint32_t nA = Q.q_count[R3D_Q_DRAW_SCAT], nB = Q.q_count[R3D_Q_DRAW_SRC];
  uint32_t i; bool have, is_src;
  while (pull_chunk(Q.q_count + R3D_Q_CURSOR_DRAW, nA, nB, Q.n_slots, i, have, is_src)) {
    if (!have) continue;
    const uint4 q = Q.q_draw[i];
    const uint32_t s = q.x;
    if (is_src) {
      // new phonon: direction = the drawn take-off angle, polarisation angle pi/2 for SH else 0 (phonons.hpp:193-207)
      const uint32_t ti = cdf_search(M.src_cdf + (size_t)q.z * M.n_toa, M.n_toa, M.src_guide + (size_t)q.z * M.guide_stride, M.guide_shift, q.y);
      const double4 t = load_toa(M.toa + ti);                   // sin th, cos th, sin ph, cos ph
      Q.dx[s] = t.x * t.w; Q.dy[s] = t.x * t.z; Q.dz[s] = t.y;
      if (q.z == R3D_RAY_SH) { Q.sx[s] = -t.z; Q.sy[s] = t.w; Q.sz[s] = 0.0; }                 // phi-hat
      else { Q.sx[s] = t.y * t.w; Q.sy[s] = t.y * t.z; Q.sz[s] = -t.x; }                      // theta-hat
    } else {
      const uint32_t ti = cdf_search(M.scat_cdf + (size_t)q.z * M.n_toa, M.n_toa, M.scat_guide + (size_t)q.z * M.guide_stride, M.guide_shift, q.y);
      const double4 t = load_toa(M.toa + ti);
      const uint32_t conv = q.z & 3u;
      double2 rp = make_double2(1.0, 0.0);                        // (cos, sin) of the relative polarisation angle
      if (conv == 3u) rp = __ldg(M.scat_spol + (size_t)(q.z >> 2) * M.n_toa + ti);
      v3 e3 = V(Q.dx[s], Q.dy[s], Q.dz[s]), s1 = V(Q.sx[s], Q.sy[s], Q.sz[s]);
      transform(e3, s1, t.x, t.y, t.z, t.w, rp.y, rp.x);
      Q.dx[s] = e3.x; Q.dy[s] = e3.y; Q.dz[s] = e3.z;
      Q.sx[s] = s1.x; Q.sy[s] = s1.y; Q.sz[s] = s1.z;
      Q.type[s] = (uint8_t)(conv & 1u);                       // PP,PS,SP,SS -> P,S,P,S
      T.v[R3D_CNT_SCATTERS]++;
      if (TRACE) Q.tr_scatters[s]++;
    }
  }
  T.flush_warp(Q.block_tally + (size_t)(tally_row0 + blockIdx.x) * R3D_NCOUNTERS);
}

// =====================================================================================================
// kernel C: interface.  Everything that happens at a face that is not a plain hand-over:
// collection (dataout.cpp:545-568, 103-216), free-surface / discontinuity R/T (phonons.cpp:429-476),
// Snell bending (phonons.cpp:311-405).
// =====================================================================================================
#define R3D_C_THREADS 128
#ifndef R3D_C_MINBLOCKS
#define R3D_C_MINBLOCKS 4      // 128 registers
#endif

// Phonon::Refraction_FullRT (phonons.cpp:429-476) + CellFace::GetRTBasis (media_cellface.cpp:122-149)
template <class Cell>
R3D_DEV void refraction_fullrt(const DevModel &M, const double *cells, Phonon &p, int face, bool adjoin, uint32_t other, RngAt &g) {
  const double *c = cells + (size_t)p.cell * M.cell_nparam;
  RTCoef rt;
  rt.init(Cell::normal(c, face, p.loc), p.dir);
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
  if (p.type == R3D_RAY_S) intype = rt.choose_spol(p.s1, g.next());        // DirectionOfMotion() of an S phonon is s1
  rt.get_coefs(intype);
  rt.choose(g.next());
  const bool reflected = (rt.choice == R_P || rt.choice == R_SV || rt.choice == R_SH);
  const v3 outdir = unit_else(rt.chosen_ray_dir(), V(0, 0, 1));          // mDir.Set(outdir.Theta(), outdir.Phi())
  p.type = (rt.choice == R_P || rt.choice == T_P) ? R3D_RAY_P : R3D_RAY_S;
  if (p.type == R3D_RAY_S) p.s1 = pol_from_pdom(outdir, rt.chosen_pdom());  // phonons.cpp:459-465
  else p.s1 = carry_pol(p.dir, p.s1, outdir);                                 // mPol is left as it was
  p.dir = outdir;
  if (!reflected) p.cell = other;
}

// Phonon::Refraction_Bend (phonons.cpp:311-405)
template <class Cell>
R3D_DEV void refraction_bend(const DevModel &M, const double *cells, Phonon &p, int face, uint32_t other) {
  const double *c = cells + (size_t)p.cell * M.cell_nparam;
  const double *o = cells + (size_t)other * M.cell_nparam;
  const v3 mdir = p.dir;
  const v3 fnorm = Cell::normal(c, face, p.loc);
  const v3 fpara = inplane_unit_perp(fnorm, mdir);
  const v3 fparash = cross(fnorm, fpara);
  const double veli = Cell::veloc(c, p.type, p.loc), velo = Cell::veloc(o, p.type, p.loc);
  const double sini = dot(fpara, mdir);
  double sino = (velo / veli) * sini;
  bool transfer; double coso;
  if (sino >= 1.0) { transfer = false; sino = sini; coso = -1.0 * dot(fnorm, mdir); }
  else { transfer = true; coso = sqrt(1.0 - (sino * sino)); }
  const v3 outraw = add(scal(fpara, sino), scal(fnorm, coso));
  const v3 outdir = unit_else(outraw, V(0, 0, 1));
  if (p.type != R3D_RAY_P) {
    const v3 svbasei = cross(fparash, mdir), svbaseo = cross(fparash, outraw);
    const double shcomi = dot(p.s1, fparash), svcomi = dot(p.s1, svbasei);
    p.s1 = pol_from_pdom(outdir, add(scal(fparash, shcomi), scal(svbaseo, svcomi)));
  } else {
    p.s1 = hats(outdir).th;                     // polout = 0 for P (phonons.cpp:366, 393)
  }
  p.dir = outdir;
  if (transfer) p.cell = other;
}

template <class Cell, bool TRACE>
__global__ void __launch_bounds__(R3D_C_THREADS, R3D_C_MINBLOCKS)
interface_kernel(const DevModel M, const Pool Q, const Job J, int cells_in_smem, uint32_t tally_row0) {
  extern __shared__ double4 smem_c4[];
  double4 *sph = smem_c4;                                     // [n_seis] (x, y, z, r_out^2 (1+eps))
  double *scells = reinterpret_cast<double *>(smem_c4 + M.n_seis);
  const uint32_t nA = Q.q_count[R3D_Q_FACE_P], nB = Q.q_count[R3D_Q_FACE_S];
  Tally T; T.clear();
  if (nA + nB > 0) {                                          // launches without work skip the table load
    for (uint32_t i = threadIdx.x; i < M.n_seis; i += blockDim.x) sph[i] = M.seis_sphere[i];
    if (cells_in_smem)
      for (uint32_t i = threadIdx.x; i < M.n_cells * M.cell_nparam; i += blockDim.x) scells[i] = M.cell_params[i];
  }
  __syncthreads();
  const double *cells = cells_in_smem ? scells : M.cell_params;

  uint32_t i; bool have, is_s;
  while (pull_chunk(Q.q_count + R3D_Q_CURSOR_FACE, nA, nB, Q.n_slots, i, have, is_s)) {
    uint32_t s = 0, fl = 0, other = 0, ordinal = 0;
    int face = 0;
    Phonon p;
    p.time = p.pathlen = p.recent = p.aexp = 0; p.loc = V(0, 0, 0); p.dir = V(0, 0, 1); p.s1 = V(1, 0, 0); p.moves = 0; p.cell = 0; p.type = 0;
    if (have) {
      const uint2 q = Q.q_face[i];
      s = q.x; face = (int)q.y;
      p.time = Q.time[s]; p.aexp = Q.aexp[s];
      p.loc = V(Q.lx[s], Q.ly[s], Q.lz[s]);
      p.dir = V(Q.dx[s], Q.dy[s], Q.dz[s]);
      p.s1 = V(Q.sx[s], Q.sy[s], Q.sz[s]);
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
          const v3 dopm = (p.type == R3D_RAY_P) ? p.dir : p.s1;   // Phonon::DirectionOfMotion (phonons.cpp:201-211)
          uint32_t bin; double e[4];
          if (seis_catch(M.seis + (size_t)k2 * R3D_SEIS_NPARAM, M.bin_dt, M.n_bins, p.time, p.loc, p.dir, dopm, p.type, exp(-p.aexp), vel, bin, e)) {
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
        Q.dx[s] = p.dir.x; Q.dy[s] = p.dir.y; Q.dz[s] = p.dir.z;
        Q.sx[s] = p.s1.x; Q.sy[s] = p.s1.y; Q.sz[s] = p.s1.z;
        Q.cell[s] = p.cell; Q.type[s] = (uint8_t)p.type; Q.ordinal[s] = g.g.ordinal;
      }
    }
  }
  T.flush_warp(Q.block_tally + (size_t)(tally_row0 + blockIdx.x) * R3D_NCOUNTERS);
}

}  // namespace r3d
