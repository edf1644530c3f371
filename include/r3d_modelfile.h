/* r3d_modelfile.h -- flat binary container for an r3d_model_desc.
 *
 * Used to move a flattened model between the host model builder, the
 * reference-side flattener (oracle/ref_harness.cpp) and the Python tests.
 * Header-only, plain C.  Layout (little-endian):
 *
 *   char     magic[8]  = "R3DMODL1"
 *   r3d_modelfile_scalars                      (fixed-size block below)
 *   13 x { uint64 nbytes; uint8 data[nbytes]; pad to 8 }  in the order
 *     toa_theta, toa_phi, src_whole_cdf, src_cdf, scat_mfp, scat_whole_cdf,
 *     scat_cdf, scat_spol, cell_params, cell_scat, face_flags,
 *     face_other_cell, seis
 */
#ifndef R3D_MODELFILE_H_
#define R3D_MODELFILE_H_

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "r3d_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct r3d_modelfile_scalars {
  double   freq_hz, ttl, bin_dt;
  double   earth_center[3];
  double   min_theta, max_theta, slow_concern;
  double   src_loc[3];
  double   cyl_radius2;
  uint64_t loop_concern;
  uint32_t n_bins;
  int32_t  ecs_radial;
  int32_t  no_deflect;
  uint32_t n_toa;
  uint32_t src_cell;
  uint32_t n_scat;
  uint32_t n_cells;
  uint32_t cell_kind;
  uint32_t cell_nparam;
  uint32_t faces_per_cell;
  uint32_t n_seis;
  uint32_t pad;
} r3d_modelfile_scalars;

#define R3D_MODELFILE_NARRAYS 13

static inline void r3d_modelfile_array_sizes(const r3d_model_desc *d, uint64_t nb[R3D_MODELFILE_NARRAYS]) {
  uint64_t nt = d->n_toa, ns = d->n_scat, nc = d->n_cells, nf = d->faces_per_cell;
  nb[0]  = nt * 8;                 /* toa_theta       */
  nb[1]  = nt * 8;                 /* toa_phi         */
  nb[2]  = 3 * 8;                  /* src_whole_cdf   */
  nb[3]  = 3 * nt * 8;             /* src_cdf         */
  nb[4]  = ns * 2 * 8;             /* scat_mfp        */
  nb[5]  = ns * 8 * 8;             /* scat_whole_cdf  */
  nb[6]  = ns * 4 * nt * 8;        /* scat_cdf        */
  nb[7]  = ns * nt * 8;            /* scat_spol       */
  nb[8]  = nc * d->cell_nparam * 8;/* cell_params     */
  nb[9]  = nc * 4;                 /* cell_scat       */
  nb[10] = nc * nf;                /* face_flags      */
  nb[11] = nc * nf * 4;            /* face_other_cell */
  nb[12] = (uint64_t)d->n_seis * R3D_SEIS_NPARAM * 8; /* seis */
}

static inline int r3d_modelfile_write(const char *path, const r3d_model_desc *d) {
  FILE *f = fopen(path, "wb");
  if (!f) return -1;
  r3d_modelfile_scalars s;
  memset(&s, 0, sizeof s);
  s.freq_hz = d->freq_hz; s.ttl = d->ttl; s.bin_dt = d->bin_dt;
  memcpy(s.earth_center, d->earth_center, sizeof s.earth_center);
  s.min_theta = d->min_theta; s.max_theta = d->max_theta; s.slow_concern = d->slow_concern;
  memcpy(s.src_loc, d->src_loc, sizeof s.src_loc);
  s.cyl_radius2 = d->cyl_radius2; s.loop_concern = d->loop_concern;
  s.n_bins = d->n_bins; s.ecs_radial = d->ecs_radial; s.no_deflect = d->no_deflect;
  s.n_toa = d->n_toa; s.src_cell = d->src_cell; s.n_scat = d->n_scat;
  s.n_cells = d->n_cells; s.cell_kind = d->cell_kind; s.cell_nparam = d->cell_nparam;
  s.faces_per_cell = d->faces_per_cell; s.n_seis = d->n_seis;
  const void *arr[R3D_MODELFILE_NARRAYS] = {
    d->toa_theta, d->toa_phi, d->src_whole_cdf, d->src_cdf, d->scat_mfp,
    d->scat_whole_cdf, d->scat_cdf, d->scat_spol, d->cell_params, d->cell_scat,
    d->face_flags, d->face_other_cell, d->seis };
  uint64_t nb[R3D_MODELFILE_NARRAYS];
  r3d_modelfile_array_sizes(d, nb);
  int ok = fwrite("R3DMODL1", 1, 8, f) == 8 && fwrite(&s, sizeof s, 1, f) == 1;
  static const char zeros[8] = {0};
  for (int i = 0; ok && i < R3D_MODELFILE_NARRAYS; i++) {
    ok = fwrite(&nb[i], 8, 1, f) == 1;
    if (ok && nb[i]) ok = fwrite(arr[i], 1, nb[i], f) == nb[i];
    uint64_t pad = (8 - nb[i] % 8) % 8;
    if (ok && pad) ok = fwrite(zeros, 1, pad, f) == pad;
  }
  if (fclose(f) != 0) ok = 0;
  return ok ? 0 : -1;
}

/* Reads a model file.  All arrays live in one malloc'd block returned in
 * *storage (free() it when done with the descriptor). */
static inline int r3d_modelfile_read(const char *path, r3d_model_desc *d, void **storage) {
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  char magic[8];
  r3d_modelfile_scalars s;
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, "R3DMODL1", 8) != 0 ||
      fread(&s, sizeof s, 1, f) != 1) { fclose(f); return -1; }
  memset(d, 0, sizeof *d);
  d->freq_hz = s.freq_hz; d->ttl = s.ttl; d->bin_dt = s.bin_dt;
  memcpy(d->earth_center, s.earth_center, sizeof s.earth_center);
  d->min_theta = s.min_theta; d->max_theta = s.max_theta; d->slow_concern = s.slow_concern;
  memcpy(d->src_loc, s.src_loc, sizeof s.src_loc);
  d->cyl_radius2 = s.cyl_radius2; d->loop_concern = s.loop_concern;
  d->n_bins = s.n_bins; d->ecs_radial = s.ecs_radial; d->no_deflect = s.no_deflect;
  d->n_toa = s.n_toa; d->src_cell = s.src_cell; d->n_scat = s.n_scat;
  d->n_cells = s.n_cells; d->cell_kind = s.cell_kind; d->cell_nparam = s.cell_nparam;
  d->faces_per_cell = s.faces_per_cell; d->n_seis = s.n_seis;
  uint64_t nb[R3D_MODELFILE_NARRAYS], total = 0;
  r3d_modelfile_array_sizes(d, nb);
  for (int i = 0; i < R3D_MODELFILE_NARRAYS; i++) total += (nb[i] + 7) / 8 * 8;
  char *blk = (char *)malloc(total ? total : 8);
  if (!blk) { fclose(f); return -1; }
  const void *arr[R3D_MODELFILE_NARRAYS];
  uint64_t off = 0;
  for (int i = 0; i < R3D_MODELFILE_NARRAYS; i++) {
    uint64_t n, padded = (nb[i] + 7) / 8 * 8;
    if (fread(&n, 8, 1, f) != 1 || n != nb[i] ||
        (padded && fread(blk + off, 1, padded, f) != padded)) { free(blk); fclose(f); return -1; }
    arr[i] = blk + off;
    off += padded;
  }
  fclose(f);
  d->toa_theta = (const double *)arr[0];  d->toa_phi = (const double *)arr[1];
  d->src_whole_cdf = (const double *)arr[2]; d->src_cdf = (const double *)arr[3];
  d->scat_mfp = (const double *)arr[4]; d->scat_whole_cdf = (const double *)arr[5];
  d->scat_cdf = (const double *)arr[6]; d->scat_spol = (const double *)arr[7];
  d->cell_params = (const double *)arr[8]; d->cell_scat = (const uint32_t *)arr[9];
  d->face_flags = (const uint8_t *)arr[10]; d->face_other_cell = (const uint32_t *)arr[11];
  d->seis = (const double *)arr[12];
  *storage = blk;
  return 0;
}

#ifdef __cplusplus
}
#endif
#endif /* R3D_MODELFILE_H_ */
