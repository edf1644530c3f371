// r3d_scatterer_gpu.cpp -- the model-build hot spot of the reference on the GPU.
//
// This file DEFINES Scatterer::PopulateProbDists() for the drop-in program.  The reference's own definition
// (scatterers.cpp:134-161: nTOA calls of ScatterParams::GSATO, ~2 s per scatterer at take-off-angle degree 9, up to 28
// scatterers per model) is compiled from the unmodified reference source and then made a weak symbol (integration/Makefile),
// so that the constructor in that translation unit reaches this definition instead.  Everything around it stays the
// reference's: the de-duplication list, PopulateWholeProbs, ComputeMFPs, ComputeDipoles and ProbDist::Integrate - the
// running sums that decide table indices are formed by the reference's own code, in its order, on the host.
//
// The G values of all take-off angles come from r3d_scatterer_g_values() (libr3dgpu.so: one thread per angle, Sato &
// Fehler 4.50-4.52 with the von Karman PSDF; parity with the reference's GSATO at 1e-10: tests/test_scatterer_tables.py).
// R3D_GPU_SCATTERERS=0 evaluates them with the reference's own ScatterParams::GSATO instead (model builds on a machine
// without a GPU; never used for the propagate path itself, which has no CPU form in this program).
#include <iostream>
#include <fstream>
#include <sstream>
#include <vector>
#include <map>
#include <string>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <complex>
#include <stdexcept>

#define private public
#define protected public
#include "scatterers.hpp"
#include "probability.hpp"
#undef private
#undef protected

#include "r3d_gpu.h"
#include "r3d_cli.hpp"

namespace {
r3d_toa_set * g_toa = 0;
const S2::S2Set * g_toa_of = 0;
size_t g_toa_n = 0;
}

void Scatterer::PopulateProbDists(ScatterParams par) {
  S2::S2Set & toa = (*pTOA);
  const size_t n = (size_t)nTOA;
  m_spol.clear();
  m_spol.resize(n);
  std::vector<double> g(4 * n);

  const char * off = getenv("R3D_GPU_SCATTERERS");
  if (off && atoi(off) == 0) {
    for (size_t k = 0; k < n; k++) {
      Real spolv;
      par.GSATO(toa[k], g[k], g[n + k], g[2 * n + k], g[3 * n + k], spolv);
      m_spol[k] = spolv;
    }
  } else {
    if (g_toa_of != pTOA || g_toa_n != n) {          // the take-off angles go to the device once per process
      if (g_toa) { r3d_toa_destroy(g_toa); g_toa = 0; }
      std::vector<double> th(n), ph(n);
      for (size_t k = 0; k < n; k++) { th[k] = toa[k].Theta(); ph[k] = toa[k].Phi(); }
      if (r3d_toa_create(th.data(), ph.data(), (uint32_t)n, r3d_cli::first_device(), &g_toa) != 0)
        throw Runtime(std::string("GPU scatterer tables: ") + r3d_last_error());
      g_toa_of = pTOA; g_toa_n = n;
    }
    r3d_scatter_params P;
    P.nu = par.nu; P.eps = par.eps; P.a = par.a; P.kappa = par.kappa; P.el = par.el; P.gam0 = par.gam0;
    if (r3d_scatterer_g_values(g_toa, &P, g.data(), m_spol.data()) != 0)
      throw Runtime(std::string("GPU scatterer tables: ") + r3d_last_error());
  }

  for (int c = 0; c < 4; c++) {
    ProbDist & D = mPDists[c];                       // (GPP, GPS, GSP, GSS = 0..3, scatterers.hpp)
    const double * gc = g.data() + (size_t)c * n;
    for (size_t k = 0; k < n; k++)
      if (gc[k] < 0.0) { D.SetRelativeProb(k, gc[k]); }     // lets the reference complain as it would (probability.hpp:128)
    D.SetRelativeProb(0, gc[0]);                     // allocates with the reference's checks (late allocation)
    memcpy(D.mDist.data(), gc, n * sizeof(double));
  }
}
