// r3d_flatten.hpp -- walk a constructed reference Model into the flat arrays of r3d_model_desc.
//
// This is the reference-side half of the drop-in boundary: it is compiled against the
// UNMODIFIED reference headers (never copied into this repository) and reads the objects the
// reference's own Model constructor built (model.cpp:220-501): cells and faces (media.hpp,
// media_cellface.hpp), the scatterer de-duplication list (scatterers.cpp:45-91), the event source
// (events.cpp:42-107), the take-off-angle set and the seismometers (dataout.cpp:42-71).
//
// The including translation unit must have included the reference headers model.hpp, media.hpp,
// phonons.hpp, scatterers.hpp, events.hpp, ecs.hpp and dataout.hpp with private members reachable
// (the usual `#define private public` after the standard headers; a maintainer would add accessor
// methods or a friend declaration instead -- see INTEGRATION.md), and r3d_gpu.h.
#ifndef R3D_FLATTEN_HPP_
#define R3D_FLATTEN_HPP_
#include <vector>
#include <map>
#include <memory>
#include <thread>
#include <stdexcept>
#include <cstring>
#include <stdint.h>

struct FlatModel {
  r3d_model_desc d;
  std::vector<double> toa_theta, toa_phi, src_whole, src_cdf, mfp, swhole, cparams, seis;
  // the scatterer tables ([n_scat][4][n_toa] and [n_scat][n_toa]: 3.5 GB for the Lop Nor model at TOA degree 9) are
  // gathered from the reference's per-table vectors by several threads into memory that is not value-initialised first
  // (as std::vector members grown by insert() they took 5.3 s of a 6.7 s RunSimulation)
  std::unique_ptr<double[]> scdf, spol;
  std::vector<uint32_t> cell_scat, other;
  std::vector<uint8_t> flags;
  std::map<const MediumCell*, uint32_t> cell_index;
};

static void put3(std::vector<double> & v, const R3::XYZ & p) {
  v.push_back(p.x()); v.push_back(p.y()); v.push_back(p.z());
}

static void Flatten(Model & Mod, FlatModel & F) {
  r3d_model_desc & d = F.d;
  memset(&d, 0, sizeof d);
  d.freq_hz = MediumCell::cmPhononFreq;
  d.ttl = Phonon::cm_ttl;
  d.bin_dt = Seismometer::cmTimePerBin;
  d.n_bins = Seismometer::cmNumBins;
  d.ecs_radial = ECS.CurvedCoords() ? 1 : 0;
  if (d.ecs_radial) {
    R3::XYZ c = ECS.GetEarthCenter();
    d.earth_center[0] = c.x(); d.earth_center[1] = c.y(); d.earth_center[2] = c.z();
  }
  d.min_theta = Phonon::cm_min_theta;
  d.max_theta = Phonon::cm_max_theta;
  d.slow_concern = Phonon::cm_slow_concern;
  d.loop_concern = Phonon::cm_loop_concern;
  d.no_deflect = Scatterer::cm_NoDeflect_b ? 1 : 0;

  // TOA set
  S2::S2Set & toa = *PhononSource::pTOA;
  d.n_toa = toa.size();
  for (size_t i = 0; i < toa.size(); i++) {
    F.toa_theta.push_back(toa[i].Theta());
    F.toa_phi.push_back(toa[i].Phi());
  }

  // cells
  const std::vector<MediumCell*> & cells = Mod.mCellArray;
  d.n_cells = cells.size();
  for (uint32_t i = 0; i < cells.size(); i++) F.cell_index[cells[i]] = i;

  // scatterers: walk the de-duplication list (scatterers.cpp:45-91)
  std::map<const Scatterer*, uint32_t> scat_index;
  std::vector<Scatterer*> scats;
  for (Scatterer * s = Scatterer::cm_ll_first; s != 0; s = s->mpllNext) {
    uint32_t idx = scat_index.size();
    scat_index[s] = idx;
    scats.push_back(s);
    F.mfp.push_back(s->mMeanFreeP[RAY_P]);
    F.mfp.push_back(s->mMeanFreeP[RAY_S]);
    for (int in = 0; in < 2; in++) {
      s->mWholeProbs[in].GetMagnitude();      // forces cumulative form
      for (int k = 0; k < 4; k++) F.swhole.push_back(s->mWholeProbs[in].mDist[k]);
    }
  }
  d.n_scat = scat_index.size();
  {
    const size_t nt = d.n_toa, ns = scats.size();
    F.scdf.reset(new double[ns * 4 * nt + 1]);
    F.spol.reset(new double[ns * nt + 1]);
    double * scdf = F.scdf.get(), * spol = F.spol.get();
    std::vector<int> bad(ns * 5, 0);
    auto gather = [&](size_t job) {            // job = scatterer * 5 + table (4 = the S->S polarisation angles)
      Scatterer * s = scats[job / 5];
      const size_t c = job % 5;
      if (c < 4) {
        s->mPDists[c].GetMagnitude();          // forces cumulative form (each ProbDist only touches its own members)
        if (s->mPDists[c].mDist.size() != nt) { bad[job] = 1; return; }
        memcpy(scdf + ((job / 5) * 4 + c) * nt, s->mPDists[c].mDist.data(), nt * sizeof(double));
      } else {
        if (s->m_spol.size() != nt) { bad[job] = 1; return; }
        memcpy(spol + (job / 5) * nt, s->m_spol.data(), nt * sizeof(double));
      }
    };
    const size_t jobs = ns * 5;
    size_t nthr = std::thread::hardware_concurrency();
    if (nthr > 16) nthr = 16;
    if (nthr < 1 || jobs * nt < (1u << 22)) nthr = 1;
    std::vector<std::thread> pool;
    for (size_t t = 1; t < nthr; t++) pool.emplace_back([&, t] { for (size_t j = t; j < jobs; j += nthr) gather(j); });
    for (size_t j = 0; j < jobs; j += nthr) gather(j);
    for (size_t t = 0; t < pool.size(); t++) pool[t].join();
    for (size_t j = 0; j < jobs; j++) if (bad[j]) throw std::runtime_error("scatterer table size differs from the take-off-angle set");
  }

  // cell records
  if (cells.empty()) throw std::runtime_error("model has no cells");
  if (dynamic_cast<RCUCylinder*>(cells[0])) {
    d.cell_kind = R3D_CELL_CYLINDER; d.cell_nparam = R3D_CYL_NPARAM; d.faces_per_cell = R3D_CYL_NFACES;
    d.cyl_radius2 = RCUCylinder::cmLossFace.mRad2;
  } else if (dynamic_cast<Tetra*>(cells[0])) {
    d.cell_kind = R3D_CELL_TETRA; d.cell_nparam = R3D_TETRA_NPARAM; d.faces_per_cell = R3D_TETRA_NFACES;
  } else if (dynamic_cast<SphereShell*>(cells[0])) {
    d.cell_kind = R3D_CELL_SHELL; d.cell_nparam = R3D_SHELL_NPARAM; d.faces_per_cell = R3D_SHELL_NFACES;
  } else throw std::runtime_error("unknown cell class");

  for (uint32_t i = 0; i < cells.size(); i++) {
    MediumCell * mc = cells[i];
    F.cell_scat.push_back(scat_index.at(mc->GetActiveScatterer()));
    std::vector<double> & p = F.cparams;
    if (d.cell_kind == R3D_CELL_CYLINDER) {
      RCUCylinder * c = dynamic_cast<RCUCylinder*>(mc);
      p.push_back(c->mVelTop[RAY_P]); p.push_back(c->mVelTop[RAY_S]);
      p.push_back(c->mDensity);
      p.push_back(c->mQ[RAY_P]); p.push_back(c->mQ[RAY_S]);
      put3(p, c->mTopFace.mNormal); put3(p, c->mTopFace.mPoint);
      put3(p, c->mBottomFace.mNormal); put3(p, c->mBottomFace.mPoint);
    } else if (d.cell_kind == R3D_CELL_SHELL) {
      SphereShell * c = dynamic_cast<SphereShell*>(mc);
      p.push_back(c->mVelCoefA[0]); p.push_back(c->mVelCoefA[1]);
      p.push_back(c->mVelCoefC[0]); p.push_back(c->mVelCoefC[1]);
      p.push_back(c->mZeroRadius2[0]); p.push_back(c->mZeroRadius2[1]);
      p.push_back(c->mDensCoefA); p.push_back(c->mDensCoefC);
      p.push_back(c->mQ[0]); p.push_back(c->mQ[1]);
      p.push_back(c->mFaces[0].mRadius); p.push_back(c->mFaces[1].mRadius);
      p.push_back(c->mFaces[0].mRad2); p.push_back(c->mFaces[1].mRad2);
    } else {
      Tetra * c = dynamic_cast<Tetra*>(mc);
      put3(p, c->mVelGrad[0]); put3(p, c->mVelGrad[1]);
      p.push_back(c->mVel0[0]); p.push_back(c->mVel0[1]);
      put3(p, c->mDensGrad); p.push_back(c->mDens0);
      p.push_back(c->mQ[0]); p.push_back(c->mQ[1]);
      for (int f = 0; f < 4; f++) { put3(p, c->mFaces[f].mNormal); put3(p, c->mFaces[f].mPoint); }
    }
    for (uint32_t f = 0; f < d.faces_per_cell; f++) {
      CellFace & cf = mc->Face(f);
      uint8_t fl = 0;
      if (cf.IsCollectionFace()) fl |= R3D_FACE_COLLECT;
      if (cf.IsReflectionFace()) fl |= R3D_FACE_REFLECT;
      if (cf.HasNeighbor())      fl |= R3D_FACE_ADJOIN;
      if (cf.GridDiscontinuity()) fl |= R3D_FACE_DISCON;
      F.flags.push_back(fl);
      F.other.push_back(cf.HasNeighbor() ? F.cell_index.at(&cf.OtherCell()) : 0xFFFFFFFFu);
    }
  }

  // source
  ShearDislocation * src = Mod.mpEventSource;
  d.src_loc[0] = src->mLoc.x(); d.src_loc[1] = src->mLoc.y(); d.src_loc[2] = src->mLoc.z();
  if (src->mpCell == 0) throw std::runtime_error("event source is not inside any cell");
  d.src_cell = F.cell_index.at(src->mpCell);
  src->mWholeProbs[0].GetMagnitude();
  for (int k = 0; k < 3; k++) F.src_whole.push_back(src->mWholeProbs[0].mDist[k]);
  for (int c = 0; c < 3; c++) {
    src->mPDists[c].GetMagnitude();
    F.src_cdf.insert(F.src_cdf.end(), src->mPDists[c].mDist.begin(), src->mPDists[c].mDist.end());
  }

  // seismometers
  d.n_seis = dataout.mSeismometers.size();
  for (size_t i = 0; i < dataout.mSeismometers.size(); i++) {
    Seismometer & s = *dataout.mSeismometers[i];
    put3(F.seis, s.mLoc); put3(F.seis, s.mAxesX1); put3(F.seis, s.mAxesX2); put3(F.seis, s.mAxesX3);
    F.seis.push_back(s.mRadiusI[0]); F.seis.push_back(s.mRadiusI[1]);
    F.seis.push_back(s.mRadiusO[0]); F.seis.push_back(s.mRadiusO[1]);
    F.seis.push_back(s.mArea[0]); F.seis.push_back(s.mArea[1]);
  }

  d.toa_theta = F.toa_theta.data(); d.toa_phi = F.toa_phi.data();
  d.src_whole_cdf = F.src_whole.data(); d.src_cdf = F.src_cdf.data();
  d.scat_mfp = F.mfp.data(); d.scat_whole_cdf = F.swhole.data();
  d.scat_cdf = F.scdf.get(); d.scat_spol = F.spol.get();
  d.cell_params = F.cparams.data(); d.cell_scat = F.cell_scat.data();
  d.face_flags = F.flags.data(); d.face_other_cell = F.other.data();
  d.seis = F.seis.data();
}

#endif  // R3D_FLATTEN_HPP_
