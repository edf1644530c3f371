// r3d_run_simulation_gpu.cpp -- the reference-side binding of the GPU propagate path.
//
// This file DEFINES Model::RunSimulation() for the reference program.  It is linked with the reference's
// own, unmodified translation units (compiled from the reference checkout where it lies; model.cpp is
// compiled with -DRunSimulation=RunSimulation_reference_cpu so that its CPU loop keeps existing under
// another name) and with libr3dgpu.so.  Everything before the loop -- command line, user_*_inc.cpp model
// plugins, grid, cells, scatterer tables, source, seismometers -- and everything after it -- the stdout
// summary, seis_traces_asc.dat and seis_NNN.octv writers (dataout.cpp:249-406, 623-694) -- is the
// reference's own code.  Only the body of the loop at model.cpp:611-625 is replaced:
//
//     for (i < mNumPhonons) { Phonon P = mpEventSource->GenerateEventPhonon(); P.Propagate(); }
//  ==>
//     flatten model -> r3d_create -> r3d_run (10 slices, for the progress lines) -> r3d_fetch
//     -> write bins / counters back into Seismometer / DataReporter objects
//
// Event reports (--reports=..., dataout.cpp:484-617): when any kind other than INV is switched on, the phonons are traced
// in batches through r3d_trace_events() and every record is printed by the reference's own DataReporter::Report* methods
// (hence its own output_phonon_dataline(): same line layout; the SID column holds the phonon index, and events come out
// phonon by phonon as in the reference).  Their side effects - the seismometer scan of ReportPhononCollected and the
// loss counters - are switched off around the printing, because the GPU already did both.
//
// Options (r3d_cli.hpp, parsed by r3d_main.cpp before the reference's parser; each also has the environment form of round 1):
//   --gpu-devices=0,1,...     CUDA devices to shard the phonon index range over   (default: 0)
//   --seed=<u64>              Philox seed (default: time(NULL), like the reference's srand(time(NULL)))
//   --num-phonons=<u64>       64-bit phonon count with the reference's K / M / B suffixes (the reference's is an int)
//   --gpu-checkpoint=<path>   after every tenth of the run write bins + counters + the phonon-index watermark to <path>;
//                             if <path> exists and belongs to the same run (count, seed, bin layout), continue from its
//                             watermark.  Phonon i always uses draw stream (seed, i), so a resumed run equals an
//                             uninterrupted one (the reference can only "resume" by adding another run, combine.m).
//                             Without --seed the seed is taken from the checkpoint; a checkpoint of another run (count,
//                             seed or bin layout) is an error, never silently overwritten.
// Environment (testing / tooling):
//   R3D_GPU_STOP_AFTER=<k>    stop after k tenths (testing the checkpoint)
//   R3D_GPU_DUMP_MODEL=<path> also write the flattened model (include/r3d_modelfile.h)
//   R3D_GPU_DUMP_ONLY=1       ... and return without simulating
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
#include <ctime>
#include <climits>
#include <chrono>
#include <algorithm>

#define private public
#define protected public
#include "model.hpp"
#include "media.hpp"
#include "phonons.hpp"
#include "scatterers.hpp"
#include "events.hpp"
#include "ecs.hpp"
#include "dataout.hpp"
#undef private
#undef protected

#include "r3d_modelfile.h"
#include "r3d_flatten.hpp"
#include "r3d_cli.hpp"

namespace {
// checkpoint file: header, then energies f64, counts u64, counters u64
struct CkptHeader { char magic[8]; uint64_t nph, seed, watermark, n_seis, n_bins; };
// true: resumed.  false: no checkpoint file.  A file that is not a checkpoint of THIS run is an error (it would be overwritten).
// `seed` is taken from the header when the user gave none.
bool ckpt_read(const char * path, uint64_t nph, uint64_t & seed, bool have_seed, size_t ns, size_t nb, uint64_t & watermark,
               std::vector<double> & e, std::vector<uint64_t> & c, std::vector<uint64_t> & k) {
  FILE * f = fopen(path, "rb");
  if (!f) return false;
  CkptHeader hd;
  if (fread(&hd, sizeof hd, 1, f) != 1 || memcmp(hd.magic, "R3DCKPT1", 8) != 0) {
    fclose(f);
    throw Runtime(std::string("checkpoint file ") + path + " exists but is not an r3d checkpoint; remove it or choose another path");
  }
  if (!have_seed) seed = hd.seed;
  if (hd.nph != nph || hd.seed != seed || hd.n_seis != ns || hd.n_bins != nb || hd.watermark > nph) {
    fclose(f);
    std::ostringstream msg;
    msg << "checkpoint file " << path << " belongs to another run (its phonons / seed / seismometers x bins: " << hd.nph << " / "
        << hd.seed << " / " << hd.n_seis << " x " << hd.n_bins << "; this run: " << nph << " / " << seed << " / " << ns << " x " << nb
        << "); remove it or choose another path";
    throw Runtime(msg.str());
  }
  bool ok = true;
  ok = ok && fread(e.data(), sizeof(double), ns * nb * R3D_BIN_NF64, f) == ns * nb * R3D_BIN_NF64;
  ok = ok && fread(c.data(), sizeof(uint64_t), ns * nb * R3D_BIN_NCNT, f) == ns * nb * R3D_BIN_NCNT;
  ok = ok && fread(k.data(), sizeof(uint64_t), R3D_NCOUNTERS, f) == R3D_NCOUNTERS;
  fclose(f);
  if (!ok) throw Runtime(std::string("checkpoint file ") + path + " is truncated; remove it");
  watermark = hd.watermark;
  return true;
}
void ckpt_write(const char * path, uint64_t nph, uint64_t seed, size_t ns, size_t nb, uint64_t watermark,
                const std::vector<double> & e, const std::vector<uint64_t> & c, const std::vector<uint64_t> & k) {
  std::string tmp = std::string(path) + ".tmp";
  FILE * f = fopen(tmp.c_str(), "wb");
  if (!f) throw Runtime(std::string("cannot write checkpoint ") + tmp);
  CkptHeader hd;
  memcpy(hd.magic, "R3DCKPT1", 8);
  hd.nph = nph; hd.seed = seed; hd.watermark = watermark; hd.n_seis = ns; hd.n_bins = nb;
  fwrite(&hd, sizeof hd, 1, f);
  fwrite(e.data(), sizeof(double), ns * nb * R3D_BIN_NF64, f);
  fwrite(c.data(), sizeof(uint64_t), ns * nb * R3D_BIN_NCNT, f);
  fwrite(k.data(), sizeof(uint64_t), R3D_NCOUNTERS, f);
  fclose(f);
  rename(tmp.c_str(), path);                 // a checkpoint is either the old one or the new one, never half of one
}
void r3d_check(int rc, const char * what) {
  if (rc != 0)   // same convention as the rest of the program: user-facing failure -> Runtime (typedefs.hpp:123)
    throw Runtime(std::string("GPU propagate path: ") + what + ": " + r3d_last_error());
}
}

// ModelParams::OutputOctaveText (model.cpp:138-...), called by the reference's main() before the model is built: the
// reference parses --num-phonons with atoi() into an int (cmdline.cpp:366-384) although ModelParams::NumPhonons is a long.
// model.cpp is compiled with -DOutputOctaveText=OutputOctaveText_reference_cpu (integration/Makefile), so this definition is
// the one main() reaches: it writes the file with the reference's own code from a copy that holds the true 64-bit count,
// so that NumPhonons in out_mparams.octv stays what vis/seisplot/combine.m:31-33 adds up.
void r3d_mparams_octave_reference(const ModelParams *, std::ostream *) asm("_ZNK11ModelParams30OutputOctaveText_reference_cpuEPSo");
void ModelParams::OutputOctaveText(std::ostream * out) const {
  ModelParams copy(*this);
  if (r3d_cli::options().have_nph) copy.NumPhonons = (long)r3d_cli::options().nph;
  r3d_mparams_octave_reference(&copy, out);
}

void Model::RunSimulation() {

  typedef std::chrono::steady_clock Clock;
  const Clock::time_point t_enter = Clock::now();
  auto seconds_since = [](Clock::time_point t) { return std::chrono::duration<double>(Clock::now() - t).count(); };

  FlatModel F;
  Flatten(*this, F);
  const double t_flatten = seconds_since(t_enter);

  if (const char * path = getenv("R3D_GPU_DUMP_MODEL")) {
    if (r3d_modelfile_write(path, &F.d) != 0) throw Runtime(std::string("cannot write model file ") + path);
    std::cerr << "r3d-gpu: wrote flattened model to " << path << "\n";
    if (getenv("R3D_GPU_DUMP_ONLY")) return;
  }

  const r3d_cli::Options & opt = r3d_cli::options();
  std::vector<int> devices = opt.devices;
  if (devices.empty()) devices.push_back(0);
  uint64_t seed = opt.have_seed ? opt.seed : (uint64_t)time(NULL);
  uint64_t nph = (mNumPhonons > 0) ? (uint64_t)mNumPhonons : 0;
  if (opt.have_nph) nph = opt.nph;

  r3d_handle * h = 0;
  const Clock::time_point t_c0 = Clock::now();
  r3d_check(r3d_create(&F.d, devices.data(), (int)devices.size(), &h), "r3d_create");
  const double t_create = seconds_since(t_c0);
  const Clock::time_point t_loop0 = Clock::now();

  std::cout << "@@ __BEGINNING_SIMULATION__" << std::endl << std::flush;

  // which event reports does the user want?  (INV alone is what waveform runs use: nothing to print unless it happens)
  uint32_t ev_mask = 0;
  if (dataout.mbReportGenerate) ev_mask |= 1u << R3D_EV_GEN;
  if (dataout.mbReportScatter)  ev_mask |= 1u << R3D_EV_SCT;
  if (dataout.mbReportCollect)  ev_mask |= 1u << R3D_EV_COL;
  if (dataout.mbReportReflect)  ev_mask |= 1u << R3D_EV_REF;
  if (dataout.mbReportTransfer) ev_mask |= 1u << R3D_EV_CEL;
  if (dataout.mbReportLost)     ev_mask |= 1u << R3D_EV_LST;
  if (dataout.mbReportTimeout)  ev_mask |= 1u << R3D_EV_TMO;
  if (dataout.mbReportInvalid)  ev_mask |= 1u << R3D_EV_INV;
  const bool reporting = (ev_mask & ~(1u << R3D_EV_INV)) != 0;

  const size_t ns = dataout.mSeismometers.size(), nb = Seismometer::cmNumBins;
  std::vector<double> e0(ns * nb * R3D_BIN_NF64 + 1, 0.0);               // what a resumed run starts from
  std::vector<uint64_t> c0(ns * nb * R3D_BIN_NCNT + 1, 0), k0(R3D_NCOUNTERS, 0);

  // ten slices so that the progress lines of model.cpp:616-628 keep appearing
  double device_seconds = 0;
  if (reporting) {
    // video runs: tens of thousands of phonons, every event printed
    std::vector<Seismometer*> seis_keep;
    seis_keep.swap(dataout.mSeismometers);                 // ReportPhononCollected must only print
    const unsigned long keep_lost = dataout.mNumLost, keep_tmo = dataout.mNumTimeout, keep_inv = dataout.mNumInvalid;
    const unsigned keep_diag = dataout.mDiagInvalid;
    const uint64_t batch = 4096;
    std::vector<r3d_event> ev(batch * 64);
    bool retraced = false;
    Phonon P(S2::ThetaPhi(0, 0), RAY_P);
    for (uint64_t lo = 0; lo < nph; lo += batch) {
      const uint64_t n = std::min<uint64_t>(batch, nph - lo);
      uint64_t got = 0;
      for (;;) {
        int rc = r3d_trace_events(h, lo, n, seed, ev_mask, ev.data(), ev.size(), &got);
        if (rc != 0) { std::string msg = r3d_last_error(); r3d_destroy(h); throw Runtime("GPU propagate path: " + msg); }
        if (got <= ev.size()) break;
        // More events than room (the whole-Earth model reports ~220 per phonon at its scripted time to live): make room and
        // trace the batch again.  Every trace also accumulates into the device bins; the bins of a report run therefore
        // come from a separate plain pass over the same phonons after the printing (below), not from these traces.
        ev.resize(got + got / 8 + 1024);
        retraced = true;
      }
      for (uint64_t i = 0; i < got; i++) {
        const r3d_event & e = ev[i];
        P.mSID = e.phonon;
        P.mType = (e.type == R3D_RAY_P) ? RAY_P : RAY_S;
        P.mTimeAlive = e.time; P.mPathLength = e.pathlen; P.mAmplitude = e.amp;
        P.mLoc = R3::XYZ(e.loc[0], e.loc[1], e.loc[2]);
        P.mDir = S2::ThetaPhi(e.theta, e.phi);
        P.mpCell = mCellArray[e.cell];
        P.mMoveCount = e.moves;
        switch (e.kind) {
          case R3D_EV_GEN: dataout.ReportNewEventPhonon(P); break;
          case R3D_EV_SCT: dataout.ReportScatterEvent(P); break;
          case R3D_EV_COL: dataout.ReportPhononCollected(P); break;
          case R3D_EV_REF: dataout.ReportReflection(P); break;
          case R3D_EV_CEL: dataout.ReportCellToCell(P); break;
          case R3D_EV_LST: dataout.ReportLostPhonon(P); break;
          case R3D_EV_TMO: dataout.ReportPhononTimeout(P); break;
          default: {       // the reason only feeds mDiagInvalid (dataout.cpp:611-617)
            int why = 0;
            while (why < 6 && !((e.reason >> why) & 1u)) why++;
            dataout.ReportInvalidPhonon(P, (DataReporter::invalid_reason_e)why);
            break;
          }
        }
      }
      std::cerr << (100 * (lo + n)) / nph << "% of " << nph << " have been cast.\n";
    }
    if (retraced) {         // a batch was traced twice: take bins and counters from one plain pass (same phonons, same draws)
      int rc = r3d_reset(h);
      if (rc == 0) rc = r3d_run(h, 0, nph, seed);
      if (rc == 0) rc = r3d_sync(h, 0);
      if (rc != 0) { std::string msg = r3d_last_error(); r3d_destroy(h); throw Runtime("GPU propagate path: " + msg); }
    }
    seis_keep.swap(dataout.mSeismometers);
    dataout.mNumLost = keep_lost; dataout.mNumTimeout = keep_tmo; dataout.mNumInvalid = keep_inv; dataout.mDiagInvalid = keep_diag;
  } else {
  const char * ckpt = opt.checkpoint.empty() ? 0 : opt.checkpoint.c_str();
  const int stop_after = getenv("R3D_GPU_STOP_AFTER") ? atoi(getenv("R3D_GPU_STOP_AFTER")) : 10;
  uint64_t watermark = 0;
  if (ckpt && ckpt_read(ckpt, nph, seed, opt.have_seed, ns, nb, watermark, e0, c0, k0))
    std::cerr << "r3d-gpu: resuming from checkpoint " << ckpt << " at phonon " << watermark << "\n";
  else { std::fill(e0.begin(), e0.end(), 0.0); std::fill(c0.begin(), c0.end(), 0); std::fill(k0.begin(), k0.end(), 0); watermark = 0; }
  // Slices: every launch has a ramp-up and a tail (its last, longest-lived phonons), worth about 20 phonons per slot of
  // the kernel, so a slice should hold a few hundred phonons per slot.  With a checkpoint file the run is cut in tenths (a
  // checkpoint after each, and the reference's "N% of ... have been cast" lines, model.cpp:616-628); without one in as
  // many tenths as leave 2e8 phonons to a slice (Lop Nor, 1e8 phonons: ten launches took 0.87 s, one takes 0.51 s).
  const int n_slices = ckpt ? 10 : (int)std::max<uint64_t>(1, std::min<uint64_t>(10, nph / 200000000ull));
  for (int slice = 0; slice < n_slices; slice++) {
    uint64_t lo = nph / n_slices * slice + (nph % n_slices) * slice / n_slices;
    uint64_t hi = nph / n_slices * (slice + 1) + (nph % n_slices) * (slice + 1) / n_slices;
    std::cerr << (100 * slice) / n_slices << "% of " << nph << " have been cast.\n";
    if (hi <= watermark) continue;             // done before the interruption
    if (lo < watermark) lo = watermark;
    int rc = r3d_run(h, lo, hi - lo, seed);
    double t = 0;
    if (rc == 0) rc = r3d_sync(h, &t);
    if (rc != 0) { std::string msg = r3d_last_error(); r3d_destroy(h); throw Runtime("GPU propagate path: " + msg); }
    device_seconds += t;
    if (ckpt) {
      std::vector<double> e(e0.size());
      std::vector<uint64_t> c(c0.size()), k(R3D_NCOUNTERS);
      uint32_t dg = 0;
      r3d_check(r3d_fetch(h, e.data(), c.data(), k.data(), &dg), "r3d_fetch");
      for (size_t i = 0; i < e.size(); i++) e[i] += e0[i];
      for (size_t i = 0; i < c.size(); i++) c[i] += c0[i];
      for (int i = 0; i < R3D_NCOUNTERS; i++) { if (i == R3D_CNT_DIAG) k[i] |= k0[i]; else k[i] += k0[i]; }
      ckpt_write(ckpt, nph, seed, ns, nb, hi, e, c, k);
    }
    if (slice + 1 >= stop_after && slice + 1 < n_slices) {
      std::cerr << "r3d-gpu: stopping after " << (slice + 1) << " tenths as asked (R3D_GPU_STOP_AFTER)\n";
      r3d_destroy(h);
      exit(3);
    }
  }
  }
  if (!reporting) std::cerr << "100% of " << nph << " have been cast.\n";
  std::cout << "@@ __SIMULATION_COMPLETE__" << std::endl;
  const double t_loop = seconds_since(t_loop0);
  const Clock::time_point t_f0 = Clock::now();

  // bins and counters back into the reference's own objects
  std::vector<double> e(ns * nb * R3D_BIN_NF64 + 1);
  std::vector<uint64_t> c(ns * nb * R3D_BIN_NCNT + 1), k(R3D_NCOUNTERS);
  uint32_t diag = 0;
  int rc = r3d_fetch(h, e.data(), c.data(), k.data(), &diag);
  if (rc != 0) { std::string msg = r3d_last_error(); r3d_destroy(h); throw Runtime("GPU propagate path: " + msg); }
  r3d_destroy(h);
  for (size_t i = 0; i < e.size(); i++) e[i] += e0[i];
  for (size_t i = 0; i < c.size(); i++) c[i] += c0[i];
  for (int i = 0; i < R3D_NCOUNTERS; i++) { if (i == R3D_CNT_DIAG) k[i] |= k0[i]; else k[i] += k0[i]; }
  diag |= (uint32_t)k0[R3D_CNT_DIAG];
  bool clipped = false;
  for (size_t s = 0; s < ns; s++) {
    Seismometer & S = *dataout.mSeismometers[s];
    for (size_t b = 0; b < nb; b++) {
      const double * eb = &e[(s * nb + b) * R3D_BIN_NF64];
      const uint64_t * cb = &c[(s * nb + b) * R3D_BIN_NCNT];
      S.mTimeBins[b].mEnergyAxes[0] += eb[0];
      S.mTimeBins[b].mEnergyAxes[1] += eb[1];
      S.mTimeBins[b].mEnergyAxes[2] += eb[2];
      S.mTimeBins[b].mEnergyByType[RAY_P] += eb[3];
      S.mTimeBins[b].mEnergyByType[RAY_S] += eb[4];
      for (int t = 0; t < 2; t++) {           // the reference's counts are 32-bit (dataout.hpp:77-93)
        uint64_t v = (uint64_t)S.mTimeBins[b].mCountByType[t] + cb[t];
        if (v > UINT_MAX) { v = UINT_MAX; clipped = true; }
        S.mTimeBins[b].mCountByType[t] = (unsigned)v;
      }
    }
  }
  if (clipped) std::cerr << "r3d-gpu: WARNING: a bin count exceeded 2^32-1 and was clipped in the output files.\n";
  dataout.mNumLost += k[R3D_CNT_LOST];
  dataout.mNumTimeout += k[R3D_CNT_TIMEOUT];
  dataout.mNumInvalid += k[R3D_CNT_INVALID];
  dataout.mDiagInvalid |= diag;

  std::cerr << "r3d-gpu: " << nph << " phonons on " << devices.size() << " device(s) in " << device_seconds
            << " s of device time (" << (device_seconds > 0 ? nph / device_seconds : 0) << " phonons/s), seed " << seed << "\n";
  // host model in -> host bins out, as this program sees it (its model arrays are pageable std::vector memory)
  std::cerr << "r3d-gpu: wall clock: flatten " << t_flatten << " s, r3d_create " << t_create << " s, loop " << t_loop
            << " s, fetch + write-back " << seconds_since(t_f0) << " s; RunSimulation total " << seconds_since(t_enter) << " s\n";

  dataout.OutputPostSimSummary();

}
