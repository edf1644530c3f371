// r3d_main.cpp -- main() of the drop-in program: the GPU path's own options, then the reference's main().
//
// The program is linked with -Wl,--wrap=main (integration/Makefile): the C runtime enters __wrap_main below, which takes the
// options of r3d_cli.hpp off the command line, keeps them for Model::RunSimulation() (r3d_run_simulation_gpu.cpp) and hands
// everything else to __real_main - the reference's own, unmodified main() - so every reference option, model plugin and
// output format stays as is.
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include "r3d_cli.hpp"

extern "C" int __real_main(int argc, char * argv[]);

namespace r3d_cli {

Options & options() { static Options o; return o; }

int first_device() { return options().devices.empty() ? 0 : options().devices[0]; }

bool parse_count(const std::string & text, uint64_t & out) {
  if (text.empty()) return false;
  std::string t = text;
  uint64_t mult = 1;
  const char last = t[t.size() - 1];
  if (last == 'K') mult = 1000ull; else if (last == 'M') mult = 1000000ull; else if (last == 'B') mult = 1000000000ull;
  if (mult != 1) t.resize(t.size() - 1);
  if (t.empty()) return false;
  char * end = 0;
  if (t.find_first_of(".eE") != std::string::npos) {          // 1e10, 2.5e9
    const double v = strtod(t.c_str(), &end);
    if (*end || !(v >= 0) || v * (double)mult > 1.8e19) return false;
    out = (uint64_t)(v * (double)mult + 0.5);
    return true;
  }
  if (t[0] == '-' || t[0] == '+') return false;
  const unsigned long long v = strtoull(t.c_str(), &end, 10);
  if (*end) return false;
  if (mult > 1 && v > 0xffffffffffffffffull / mult) return false;
  out = (uint64_t)v * mult;
  return true;
}

static void parse_devices(const std::string & s, std::vector<int> & out) {
  out.clear();
  std::stringstream ss(s);
  std::string tok;
  while (std::getline(ss, tok, ',')) if (!tok.empty()) out.push_back(atoi(tok.c_str()));
}

}  // namespace r3d_cli

extern "C" int __wrap_main(int argc, char * argv[]) {
  using namespace r3d_cli;
  Options & o = options();
  // environment first (round-1 interface), options override
  if (const char * s = getenv("R3D_GPU_DEVICES")) parse_devices(s, o.devices);
  if (const char * s = getenv("R3D_GPU_SEED")) { o.seed = strtoull(s, 0, 0); o.have_seed = true; }
  if (const char * s = getenv("R3D_GPU_NUM_PHONONS")) { o.have_nph = parse_count(s, o.nph); }
  if (const char * s = getenv("R3D_GPU_CHECKPOINT")) o.checkpoint = s;

  std::vector<std::string> keep;
  keep.push_back(argc > 0 ? argv[0] : "r3d_gpu_main");
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i], name = a, value;
    bool have_value = false;
    const size_t eq = a.find('=');
    if (eq != std::string::npos && a.compare(0, 2, "--") == 0) { name = a.substr(0, eq); value = a.substr(eq + 1); have_value = true; }
    const bool ours = (name == "--gpu-devices" || name == "--seed" || name == "--gpu-checkpoint");
    const bool count = (name == "--num-phonons" || name == "-N");
    if (!ours && !count) { keep.push_back(a); continue; }
    if (!have_value) {                                   // "--opt value" form
      if (i + 1 >= argc) { fprintf(stderr, "** Error processing command-line option: %s\n** Message: Required value not provided\n** Exiting...\n", a.c_str()); return 1; }
      value = argv[++i];
    }
    if (name == "--gpu-devices") parse_devices(value, o.devices);
    else if (name == "--seed") { o.seed = strtoull(value.c_str(), 0, 0); o.have_seed = true; }
    else if (name == "--gpu-checkpoint") o.checkpoint = value;
    else {
      uint64_t n = 0;
      if (!parse_count(value, n)) { keep.push_back(name + "=" + value); continue; }      // let the reference's parser complain
      o.nph = n; o.have_nph = true;
      char buf[32];
      snprintf(buf, sizeof buf, "%llu", (unsigned long long)(n > (uint64_t)INT_MAX ? (uint64_t)INT_MAX : n));
      keep.push_back("--num-phonons=" + std::string(buf));
    }
  }
  std::vector<char *> av;
  for (size_t i = 0; i < keep.size(); i++) av.push_back(const_cast<char *>(keep[i].c_str()));
  av.push_back(0);
  return __real_main((int)keep.size(), av.data());
}
