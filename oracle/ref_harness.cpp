// ref_harness.cpp -- TEST INFRASTRUCTURE ONLY (never part of the product path).
//
// A driver around the UNMODIFIED reference objects (compiled by oracle/Makefile
// from /root/reference where they lie).  It gives the test-suite three things
// the reference does not offer by itself:
//
//   R3D_HARNESS=dump   build the Model exactly as the reference's main() would
//                      for the given command line, walk its cells / faces /
//                      scatterers / source / seismometers and write the flat
//                      r3d_model_desc (include/r3d_gpu.h) to $R3D_HARNESS_OUT.
//   R3D_HARNESS=run    as dump, then run the reference's own
//                      GenerateEventPhonon()+Propagate() loop (model.cpp:611-614)
//                      with rand() REPLACED by the Philox4x32-10 stream the GPU
//                      path uses, keyed by (seed, phonon index, draw ordinal).
//                      Writes bins, counters and per-phonon end states, so the
//                      reference and the GPU path can be compared phonon by
//                      phonon, not only statistically.
//   R3D_HARNESS=randrun  same loop but with the C library generator seeded by
//                      $R3D_HARNESS_SEED (statistical comparison runs).
//   R3D_HARNESS=vectors  deterministic sub-kernel golden vectors.
//   R3D_HARNESS_MUTATE=key=value,...  (any mode) after the reference built the Model, overwrite members of ITS objects
//                      before flattening / running, so that the reference itself decides what a degenerate model does:
//                      ttl, slow, loop (Phonon::cm_ttl / cm_slow_concern / cm_loop_concern), mfp_p, mfp_s (every
//                      scatterer's mean free path), cyl_vel_p, cyl_vel_s (RCUCylinder::mVelTop), shell_c_p, shell_c_s
//                      (SphereShell::mVelCoefC), shell_zr2_p, shell_zr2_s (SphereShell::mZeroRadius2), no_reflect=1
//                      (CellFace::mReflect cleared on every reflecting face).  Used for the invalid-phonon fixtures (phonons.cpp:554-584): no
//                      stock model produces a single INV phonon.
//   R3D_HARNESS=scatparams  the ScatterParams (nu eps a kappa el gam0) of every scatterer.
//
// Private members are reached with the usual "#define private public" trick,
// applied after the standard headers are in; the reference objects themselves
// are compiled without it, only this file sees through the access specifiers.

#include <iostream>
#include <fstream>
#include <sstream>
#include <vector>
#include <map>
#include <string>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <cmath>
#include <complex>
#include <stdexcept>
#include <cassert>
#include <algorithm>
#include <iomanip>
#include <ctime>

#define private public
#define protected public
#define main r3d_reference_main_unused
#include "main.cpp"          // the reference's option parser (process_option)
#undef main
#include "media.hpp"
#include "phonons.hpp"
#include "scatterers.hpp"
#include "events.hpp"
#include "ecs.hpp"
#undef private
#undef protected

#include "r3d_modelfile.h"

// ---------------------------------------------------------------------------
// rand() interposition
// ---------------------------------------------------------------------------
namespace {

enum RandMode { RAND_LIBC, RAND_PHILOX, RAND_SCRIPT };
RandMode g_rand_mode = RAND_LIBC;
uint64_t g_seed = 0, g_phonon = 0;
uint32_t g_ordinal = 0;
std::vector<int> g_script;
size_t g_script_pos = 0;

inline void philox_round(uint32_t c[4], const uint32_t k[2]) {
  const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
  const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
  uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
  uint32_t n1 = (uint32_t)p1;
  uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
  uint32_t n3 = (uint32_t)p0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}

// Philox4x32-10; counter = (idx_lo, idx_hi, block, 0), key = (seed_lo, seed_hi)
void philox4x32_10(uint64_t seed, uint64_t idx, uint32_t block, uint32_t out[4]) {
  uint32_t c[4] = {(uint32_t)idx, (uint32_t)(idx >> 32), block, 0u};
  uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
  for (int r = 0; r < 10; r++) {
    philox_round(c, k);
    k[0] += 0x9E3779B9u;
    k[1] += 0xBB67AE85u;
  }
  for (int i = 0; i < 4; i++) out[i] = c[i];
}

} // namespace

extern "C" int rand(void) {
  switch (g_rand_mode) {
  case RAND_PHILOX: {
    uint32_t w[4];
    philox4x32_10(g_seed, g_phonon, g_ordinal >> 2, w);
    int k = (int)(w[g_ordinal & 3] >> 1);     // 31 bits: [0, RAND_MAX]
    g_ordinal++;
    return k;
  }
  case RAND_SCRIPT:
    if (g_script_pos >= g_script.size()) { std::cerr << "rand script exhausted\n"; exit(3); }
    g_ordinal++;
    return g_script[g_script_pos++];
  default:
    return (int)random();
  }
}

#include "r3d_flatten.hpp"   // shared with the product's reference-side stub (integration/)

// R3D_HARNESS_MUTATE: see the header comment
static void Mutate(Model & Mod, const char * spec) {
  std::stringstream ss(spec);
  std::string item;
  while (std::getline(ss, item, ',')) {
    size_t eq = item.find('=');
    if (eq == std::string::npos) { std::cerr << "harness: bad mutation " << item << "\n"; exit(1); }
    const std::string key = item.substr(0, eq);
    const double v = strtod(item.c_str() + eq + 1, 0);
    if (key == "ttl") Phonon::cm_ttl = v;
    else if (key == "slow") Phonon::cm_slow_concern = v;
    else if (key == "loop") Phonon::cm_loop_concern = (unsigned long)v;
    else if (key == "mfp_p" || key == "mfp_s") {
      for (Scatterer * s = Scatterer::cm_ll_first; s != 0; s = s->mpllNext) s->mMeanFreeP[key == "mfp_p" ? RAY_P : RAY_S] = v;
    } else if (key == "cyl_vel_p" || key == "cyl_vel_s") {
      for (size_t i = 0; i < Mod.mCellArray.size(); i++) {
        RCUCylinder * c = dynamic_cast<RCUCylinder*>(Mod.mCellArray[i]);
        if (!c) { std::cerr << "harness: " << key << " needs a cylinder model\n"; exit(1); }
        c->mVelTop[key == "cyl_vel_p" ? RAY_P : RAY_S] = v;
      }
    } else if (key == "shell_c_p" || key == "shell_c_s") {
      for (size_t i = 0; i < Mod.mCellArray.size(); i++) {
        SphereShell * c = dynamic_cast<SphereShell*>(Mod.mCellArray[i]);
        if (!c) { std::cerr << "harness: " << key << " needs a shell model\n"; exit(1); }
        c->mVelCoefC[key == "shell_c_p" ? RAY_P : RAY_S] = v;
      }
    } else if (key == "shell_zr2_p" || key == "shell_zr2_s") {
      for (size_t i = 0; i < Mod.mCellArray.size(); i++) {
        SphereShell * c = dynamic_cast<SphereShell*>(Mod.mCellArray[i]);
        if (!c) { std::cerr << "harness: " << key << " needs a shell model\n"; exit(1); }
        c->mZeroRadius2[key == "shell_zr2_p" ? RAY_P : RAY_S] = v;
      }
    } else if (key == "no_reflect") {        // every reflecting face stops reflecting: phonons that reach the free surface are lost
      for (size_t i = 0; i < Mod.mCellArray.size(); i++)
        for (Index f = 0; f < Mod.mCellArray[i]->NumFaces(); f++)
          if (Mod.mCellArray[i]->Face(f).IsReflectionFace()) Mod.mCellArray[i]->Face(f).SetReflect(v == 0);
    } else { std::cerr << "harness: unknown mutation " << key << "\n"; exit(1); }
  }
}

// ---------------------------------------------------------------------------
// result files
// ---------------------------------------------------------------------------
// bins file: "R3DBINS1", u32 n_seis, u32 n_bins, u64 counters[8], then
// f64 energies[n_seis][n_bins][5], u64 counts[n_seis][n_bins][2]
static void WriteBins(const std::string & path, uint64_t nphonons) {
  std::ofstream f(path.c_str(), std::ios::binary);
  uint32_t ns = dataout.mSeismometers.size(), nb = Seismometer::cmNumBins;
  uint64_t counters[R3D_NCOUNTERS] = {0};
  counters[R3D_CNT_LOST] = dataout.mNumLost;
  counters[R3D_CNT_TIMEOUT] = dataout.mNumTimeout;
  counters[R3D_CNT_INVALID] = dataout.mNumInvalid;
  counters[7] = dataout.mDiagInvalid;
  counters[6] = nphonons;
  f.write("R3DBINS1", 8);
  f.write((const char*)&ns, 4); f.write((const char*)&nb, 4);
  f.write((const char*)counters, sizeof counters);
  for (uint32_t s = 0; s < ns; s++) {
    Seismometer & S = *dataout.mSeismometers[s];
    for (uint32_t b = 0; b < nb; b++) {
      double e[5] = {S.mTimeBins[b].mEnergyAxes[0], S.mTimeBins[b].mEnergyAxes[1], S.mTimeBins[b].mEnergyAxes[2],
                     S.mTimeBins[b].mEnergyByType[0], S.mTimeBins[b].mEnergyByType[1]};
      f.write((const char*)e, sizeof e);
    }
  }
  for (uint32_t s = 0; s < ns; s++) {
    Seismometer & S = *dataout.mSeismometers[s];
    for (uint32_t b = 0; b < nb; b++) {
      uint64_t c[2] = {S.mTimeBins[b].mCountByType[0], S.mTimeBins[b].mCountByType[1]};
      f.write((const char*)c, sizeof c);
    }
  }
}

static uint64_t TotalCatches() {
  uint64_t t = 0;
  for (size_t s = 0; s < dataout.mSeismometers.size(); s++) {
    Seismometer & S = *dataout.mSeismometers[s];
    for (uint32_t b = 0; b < Seismometer::cmNumBins; b++)
      t += S.mTimeBins[b].mCountByType[0] + S.mTimeBins[b].mCountByType[1];
  }
  return t;
}

// ---------------------------------------------------------------------------
// sub-kernel golden vectors
// ---------------------------------------------------------------------------
static uint64_t g_lcg = 0x9E3779B97F4A7C15ull;
static double urand() {   // private generator for test inputs (not rand())
  g_lcg = g_lcg * 6364136223846793005ull + 1442695040888963407ull;
  return (double)(g_lcg >> 11) / 9007199254740992.0;
}
static void wr(std::ofstream & f, const double * v, int n) {
  f << std::setprecision(17);
  for (int i = 0; i < n; i++) f << (i ? " " : "") << v[i];
  f << "\n";
}

static void VectorsModelBound(Model & Mod, FlatModel & F, const std::string & dir, int ncase) {
  // (iii) GetPathToBoundary / AdvanceLength TravelRecs, (ii) face distances
  // implicitly, (vi) CatchPhonon
  std::ofstream fp((dir + "/path_to_boundary.txt").c_str());
  std::ofstream fa((dir + "/advance.txt").c_str());
  const std::vector<MediumCell*> & cells = Mod.mCellArray;
  for (int n = 0; n < ncase; n++) {
    uint32_t ci = (uint32_t)(urand() * cells.size()) % cells.size();
    MediumCell * mc = cells[ci];
    // a point inside the cell: blend of "face points"
    R3::XYZ loc;
    if (F.d.cell_kind == R3D_CELL_CYLINDER) {
      RCUCylinder * c = dynamic_cast<RCUCylinder*>(mc);
      double zt = c->mTopFace.mPoint.z(), zb = c->mBottomFace.mPoint.z();
      double r = sqrt(urand()) * 0.7 * sqrt(F.d.cyl_radius2), a = urand() * 2 * M_PI;
      double w = urand();
      if (n % 7 == 0) w = 0.0;            // exactly on the top face
      loc = R3::XYZ(r * cos(a), r * sin(a), zt + (zb - zt) * w);
    } else if (F.d.cell_kind == R3D_CELL_SHELL) {
      SphereShell * c = dynamic_cast<SphereShell*>(mc);
      double rt = c->mFaces[0].mRadius, rb = -c->mFaces[1].mRadius;
      double w = urand();
      if (n % 7 == 0) w = 0.0;
      double r = rt + (rb - rt) * w;
      double th = acos(2 * urand() - 1), ph = urand() * 2 * M_PI;
      loc = R3::XYZ(r * sin(th) * cos(ph), r * sin(th) * sin(ph), r * cos(th));
    } else {
      Tetra * c = dynamic_cast<Tetra*>(mc);
      // nodes are the face points: face f's mPoint is a node not opposite... use
      // barycentric blend of the four face points (each is a tetra vertex)
      double w[4], ws = 0;
      for (int k = 0; k < 4; k++) { w[k] = 0.05 + urand(); ws += w[k]; }
      double x = 0, y = 0, z = 0;
      for (int k = 0; k < 4; k++) {
        x += w[k] / ws * c->mFaces[k].mPoint.x();
        y += w[k] / ws * c->mFaces[k].mPoint.y();
        z += w[k] / ws * c->mFaces[k].mPoint.z();
      }
      loc = R3::XYZ(x, y, z);
      if (mc->IsPointInside(loc) > 0) { n--; continue; }
    }
    double th = acos(2 * urand() - 1), ph = (urand() * 2 - 1) * M_PI;
    if (n % 11 == 0) th = (n % 22 == 0) ? 1e-7 : M_PI - 1e-7;   // vertical rays
    if (n % 13 == 0) th = M_PI / 2;                             // grazing
    S2::ThetaPhi dir(th, ph);
    raytype rt = (urand() < 0.5) ? RAY_P : RAY_S;
    TravelRec tr = mc->GetPathToBoundary(rt, loc, dir);
    int face = -1;
    for (uint32_t f = 0; f < F.d.faces_per_cell; f++) if (&mc->Face(f) == tr.pFace) face = f;
    double in[7] = {(double)ci, (double)rt, loc.x(), loc.y(), loc.z(), th, ph};
    double out[9] = {tr.PathLength, tr.TravelTime, tr.NewLoc.x(), tr.NewLoc.y(), tr.NewLoc.z(),
                     tr.NewDir.Theta(), tr.NewDir.Phi(), tr.Attenuation, (double)face};
    wr(fp, in, 7); wr(fp, out, 9);
    double len = urand() * (std::isfinite(tr.PathLength) ? tr.PathLength : 10.0);
    TravelRec ta = mc->AdvanceLength(rt, len, loc, dir);
    double in2[8] = {(double)ci, (double)rt, loc.x(), loc.y(), loc.z(), th, ph, len};
    double out2[9] = {ta.PathLength, ta.TravelTime, ta.NewLoc.x(), ta.NewLoc.y(), ta.NewLoc.z(),
                      ta.NewDir.Theta(), ta.NewDir.Phi(), ta.Attenuation, -1.0};
    wr(fa, in2, 8); wr(fa, out2, 9);
  }
}

static void VectorsFree(const std::string & dir, int ncase) {
  // (iv) Transform
  {
    std::ofstream f((dir + "/transform.txt").c_str());
    for (int n = 0; n < ncase; n++) {
      double th = acos(2 * urand() - 1), ph = (urand() * 2 - 1) * M_PI, pol = (urand() * 2 - 1) * M_PI;
      double rth = acos(2 * urand() - 1), rph = (urand() * 2 - 1) * M_PI, rpol = (urand() * 2 - 1) * M_PI;
      if (n % 9 == 0) { rth = 1e-7; rph = 0; rpol = 0; }
      if (n % 10 == 0) { th = 1e-7; }
      Phonon P(S2::ThetaPhi(th, ph), RAY_S);
      P.SetPolarization(pol);
      Phonon R(S2::ThetaPhi(rth, rph), RAY_S);
      R.SetPolarization(rpol);
      P.Transform(R);
      double in[6] = {th, ph, pol, rth, rph, rpol};
      double out[3] = {P.mDir.Theta(), P.mDir.Phi(), P.mPol};
      wr(f, in, 6); wr(f, out, 3);
    }
  }
  // (i)+RT: RTCoef on random interfaces, incl. free surface
  {
    std::ofstream f((dir + "/rtcoef.txt").c_str());
    for (int n = 0; n < ncase; n++) {
      double nth = acos(2 * urand() - 1), nph = (urand() * 2 - 1) * M_PI;
      R3::XYZ fn(sin(nth) * cos(nph), sin(nth) * sin(nph), cos(nth));
      if (n % 5 == 0) fn = R3::XYZ(0, 0, 1);
      // a direction in the outward hemisphere of fn
      R3::XYZ dir;
      do {
        double dth = acos(2 * urand() - 1), dph = (urand() * 2 - 1) * M_PI;
        dir = R3::XYZ(sin(dth) * cos(dph), sin(dth) * sin(dph), cos(dth));
      } while (dir.Dot(fn) <= 0);
      if (n % 17 == 0) dir = fn;        // normal incidence (parallel fallback path)
      RTCoef rt(fn, dir);
      bool freesurf = (n % 3 == 0);
      rt.DensityR = 2.0 + urand() * 2; rt.VelocR[RAY_P] = 4 + 4 * urand(); rt.VelocR[RAY_S] = rt.VelocR[RAY_P] / (1.5 + 0.5 * urand());
      if (freesurf) { rt.DensityT = 0; rt.VelocT[RAY_P] = 1e-12; rt.VelocT[RAY_S] = 1e-12; rt.NoTransmit = true; }
      else { rt.DensityT = 2.0 + urand() * 2; rt.VelocT[RAY_P] = 4 + 4 * urand(); rt.VelocT[RAY_S] = rt.VelocT[RAY_P] / (1.5 + 0.5 * urand()); }
      int intype = n % 3 == 0 ? RAY_P : (n % 3 == 1 ? RAY_SH : RAY_SV);
      if (n % 4 == 0) intype = (int)(urand() * 3) % 3;
      int k = (int)(urand() * 2147483647.0);
      if (n % 23 == 0) k = 0;
      g_rand_mode = RAND_SCRIPT; g_script.assign(1, k); g_script_pos = 0;
      rt.GetCoefs((raytype)intype);
      rt.Choose();
      R3::XYZ od = rt.GetChosenRayDirection();
      R3::XYZ pd = rt.GetChosenParticleDOM();
      g_rand_mode = RAND_LIBC;
      double in[15] = {fn.x(), fn.y(), fn.z(), dir.x(), dir.y(), dir.z(), rt.DensityR, rt.VelocR[0], rt.VelocR[1],
                       rt.DensityT, rt.VelocT[0], rt.VelocT[1], (double)intype, rt.NoTransmit ? 1.0 : 0.0, (double)k};
      double out[13] = {rt.mProb[0], rt.mProb[1], rt.mProb[2], rt.mProb[3], rt.mProb[4], rt.mProb[5],
                        (double)rt.mChoice, od.x(), od.y(), od.z(), pd.x(), pd.y(), pd.z()};
      wr(f, in, 15); wr(f, out, 13);
    }
  }
  // the reference's own --rtcoef-test table (rtcoef.cpp:687-742): 3 x 100 rows of probabilities
  {
    std::ofstream f((dir + "/rtcoef_test_table.txt").c_str());
    for (int irt = 0; irt < 3; irt++) for (int isin = 0; isin < 100; isin++) {
      Real theta = Geometry::Pi90 * ((Real)isin / (100 - 1));
      S2::ThetaPhi norm(0, 0), incidence(theta, 0);
      RTCoef rt(norm, incidence);
      rt.DensityR = 10; rt.DensityT = 8; rt.VelocR[RAY_P] = 8; rt.VelocT[RAY_P] = 4; rt.VelocR[RAY_S] = 4; rt.VelocT[RAY_S] = 2;
      rt.GetCoefs((raytype)irt);
      double row[9] = {(double)irt, theta, rt.mSini, rt.mProb[0], rt.mProb[3], rt.mProb[1], rt.mProb[4], rt.mProb[2], rt.mProb[5]};
      wr(f, row, 9);
    }
  }
}

static void VectorsCdfAndCatch(Model & Mod, FlatModel & F, const std::string & dir, int ncase) {
  // (v) GetRandomIndex on the real CDFs for scripted draws incl. 0 and RAND_MAX
  {
    std::ofstream f((dir + "/cdf_search.txt").c_str());
    Scatterer * s = Scatterer::cm_ll_first;
    ShearDislocation * src = Mod.mpEventSource;
    for (int n = 0; n < ncase; n++) {
      int which = n % 9;      // 0-3 scatter PDists, 4-5 scatter whole, 6-8 source PDists
      int k = (int)(urand() * 2147483648.0);
      if (n < 9) k = 0;
      else if (n < 18) k = RAND_MAX;
      else if (n < 27) k = 1;
      g_rand_mode = RAND_SCRIPT; g_script.assign(1, k); g_script_pos = 0;
      Index idx;
      if (which < 4) idx = s->mPDists[which].GetRandomIndex();
      else if (which < 6) idx = s->mWholeProbs[which - 4].GetRandomIndex();
      else idx = src->mPDists[which - 6].GetRandomIndex();
      g_rand_mode = RAND_LIBC;
      double row[3] = {(double)which, (double)k, (double)idx};
      wr(f, row, 3);
    }
  }
  // (vi) CatchPhonon: bin index + energies for random arrivals near seismometers
  if (!dataout.mSeismometers.empty()) {
    std::ofstream f((dir + "/catch.txt").c_str());
    for (int n = 0; n < ncase; n++) {
      uint32_t si = (uint32_t)(urand() * dataout.mSeismometers.size()) % dataout.mSeismometers.size();
      Seismometer & S = *dataout.mSeismometers[si];
      raytype rt = urand() < 0.5 ? RAY_P : RAY_S;
      double ro = S.mRadiusO[rt];
      double rr = ro * 1.3 * sqrt(urand()), a = urand() * 2 * M_PI;
      // offset in the seismometer's own horizontal plane
      R3::XYZ loc = S.mLoc + S.mAxesX1.ScaledBy(rr * cos(a)) + S.mAxesX2.ScaledBy(rr * sin(a));
      double th = acos(2 * urand() - 1), ph = (urand() * 2 - 1) * M_PI, pol = (urand() * 2 - 1) * M_PI;
      Phonon P(loc, S2::ThetaPhi(th, ph), rt);
      P.SetPolarization(pol);
      P.mTimeAlive = urand() * Seismometer::cmTimePerBin * Seismometer::cmNumBins * 1.1;
      if (n % 19 == 0) P.mTimeAlive = 0.0;
      P.mAmplitude = urand();
      MediumCell * cell = Mod.FindCellContainingPoint(loc);
      if (!cell) { n--; continue; }
      P.InsertInto(cell);
      double vel = P.Velocity();
      // snapshot the seismometer, catch, diff
      std::vector<Seismometer::BinRecord> before(S.mTimeBins, S.mTimeBins + Seismometer::cmNumBins);
      S.CatchPhonon(P);
      double caught = 0, bin = -1, e[4] = {0, 0, 0, 0};
      for (uint32_t b = 0; b < Seismometer::cmNumBins; b++) {
        if (S.mTimeBins[b].mCountByType[rt] != before[b].mCountByType[rt]) {
          caught = 1; bin = b;
          e[0] = S.mTimeBins[b].mEnergyAxes[0] - before[b].mEnergyAxes[0];
          e[1] = S.mTimeBins[b].mEnergyAxes[1] - before[b].mEnergyAxes[1];
          e[2] = S.mTimeBins[b].mEnergyAxes[2] - before[b].mEnergyAxes[2];
          e[3] = S.mTimeBins[b].mEnergyByType[rt] - before[b].mEnergyByType[rt];
          S.mTimeBins[b] = before[b];   // restore so differences stay exact
        }
      }
      double in[28];
      for (int k = 0; k < 18; k++) in[k] = F.seis[si * 18 + k];
      in[18] = P.mTimeAlive; in[19] = loc.x(); in[20] = loc.y(); in[21] = loc.z();
      in[22] = th; in[23] = ph; in[24] = pol; in[25] = (double)rt; in[26] = P.mAmplitude; in[27] = vel;
      double out[6] = {caught, bin, e[0], e[1], e[2], e[3]};
      wr(f, in, 28); wr(f, out, 6);
    }
  }
}

// ---------------------------------------------------------------------------
int main(int argc, char * argv[]) {
  const char * mode_c = getenv("R3D_HARNESS");
  std::string mode = mode_c ? mode_c : "dump";
  const char * out_c = getenv("R3D_HARNESS_OUT");
  std::string out = out_c ? out_c : "r3d_harness_out";
  const char * seed_c = getenv("R3D_HARNESS_SEED");
  g_seed = seed_c ? strtoull(seed_c, 0, 0) : 20261018ull;
  const char * first_c = getenv("R3D_HARNESS_FIRST");
  uint64_t first = first_c ? strtoull(first_c, 0, 0) : 0;

  if (mode == "vectors-free") {      // no model needed
    VectorsFree(out, 400);
    return 0;
  }

  MissionParams mission;
  ModelParams MParams;
  dataout.SuppressAllReports();
  CmdOpt::OptList opt_list = CmdOpt::PackageArgCArgV(argc, argv);
  for (Index i = 0; i < opt_list.size(); i++) {
    try { process_option(opt_list[i], MParams, mission); }
    catch (std::exception & e) {
      std::cerr << "harness: bad option " << opt_list[i].GetOptionText() << ": " << e.what() << "\n";
      return 1;
    }
  }
  try {
    Model Mod(MParams);
    if (const char * mut = getenv("R3D_HARNESS_MUTATE")) Mutate(Mod, mut);
    FlatModel F;
    Flatten(Mod, F);
    if (mode == "dump") {
      if (r3d_modelfile_write(out.c_str(), &F.d) != 0) { std::cerr << "cannot write " << out << "\n"; return 2; }
      std::cerr << "harness: wrote model (" << F.d.n_cells << " cells, " << F.d.n_scat << " scatterers, "
                << F.d.n_toa << " TOA, " << F.d.n_seis << " seismometers) to " << out << "\n";
      return 0;
    }
    if (mode == "scatparams") {            // the ScatterParams of every scatterer, in the flattener's order
      std::ofstream f(out.c_str());
      f.precision(17);
      for (Scatterer * sc = Scatterer::cm_ll_first; sc != 0; sc = sc->mpllNext)
        f << sc->mParams.nu << " " << sc->mParams.eps << " " << sc->mParams.a << " " << sc->mParams.kappa << " "
          << sc->mParams.el << " " << sc->mParams.gam0 << "\n";
      return 0;
    }
    if (mode == "vectors") {
      VectorsModelBound(Mod, F, out, 300);
      VectorsCdfAndCatch(Mod, F, out, 300);
      return 0;
    }
    if (mode == "run" || mode == "randrun") {
      if (r3d_modelfile_write((out + ".model").c_str(), &F.d) != 0) return 2;
      long N = Mod.mNumPhonons;
      std::vector<r3d_phonon_final> finals;
      bool want_trace = getenv("R3D_HARNESS_TRACE") != 0;
      if (mode == "randrun") { srandom((unsigned)g_seed); g_rand_mode = RAND_LIBC; }
      clock_t t0 = clock();
      for (long i = 0; i < N; i++) {
        if (mode == "run") { g_rand_mode = RAND_PHILOX; g_phonon = first + i; }
        g_ordinal = 0;
        unsigned long l0 = dataout.mNumLost, t0c = dataout.mNumTimeout, i0 = dataout.mNumInvalid;
        unsigned d0 = dataout.mDiagInvalid;
        dataout.mDiagInvalid = 0;
        Phonon P = Mod.mpEventSource->GenerateEventPhonon();   // model.cpp:612
        P.Propagate();                                          // model.cpp:613
        if (want_trace) {
          r3d_phonon_final r;
          memset(&r, 0, sizeof r);
          r.time = P.mTimeAlive; r.pathlen = P.mPathLength; r.amp = P.mAmplitude;
          r.loc[0] = P.mLoc.x(); r.loc[1] = P.mLoc.y(); r.loc[2] = P.mLoc.z();
          r.theta = P.mDir.Theta(); r.phi = P.mDir.Phi(); r.pol = P.mPol;
          r.moves = P.mMoveCount; r.cell = F.cell_index.at(P.mpCell); r.type = P.mType;
          if (dataout.mNumLost != l0) r.fate = R3D_FATE_LOST;
          else if (dataout.mNumTimeout != t0c) r.fate = R3D_FATE_TIMEOUT;
          else if (dataout.mNumInvalid != i0) r.fate = R3D_FATE_INVALID | (dataout.mDiagInvalid << 8);
          r.draws = g_ordinal;
          r.catches = 0xFFFFFFFFu; r.scatters = 0xFFFFFFFFu; r.iters = 0xFFFFFFFFu;
          finals.push_back(r);
        }
        dataout.mDiagInvalid |= d0;
      }
      g_rand_mode = RAND_LIBC;
      double secs = (double)(clock() - t0) / CLOCKS_PER_SEC;
      WriteBins(out + ".bins", N);
      if (want_trace) {
        std::ofstream f((out + ".trace").c_str(), std::ios::binary);
        f.write((const char*)finals.data(), finals.size() * sizeof(r3d_phonon_final));
      }
      std::cerr << "harness: " << N << " phonons in " << secs << " s CPU ("
                << (N / secs) << " phonons/s), catches=" << TotalCatches()
                << " lost=" << dataout.mNumLost << " timeout=" << dataout.mNumTimeout
                << " invalid=" << dataout.mNumInvalid << "\n";
      std::cout << "HARNESS_SIM_SECONDS " << secs << "\n";
      return 0;
    }
    std::cerr << "harness: unknown mode " << mode << "\n";
    return 1;
  } catch (std::exception & e) {
    std::cerr << "harness: " << e.what() << "\n";
    return 1;
  }
}
