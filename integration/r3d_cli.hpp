// r3d_cli.hpp -- the command-line options the GPU path adds to the reference program.
//
//   --gpu-devices=0,1,...      CUDA devices to shard the phonon index range over (default 0; env R3D_GPU_DEVICES)
//   --seed=<u64>               Philox seed; a run is reproducible given (seed, phonon count) whatever the device count
//                              (default time(NULL), like the reference's srand(time(NULL)), model.cpp:235; env R3D_GPU_SEED)
//   --num-phonons=<n>, -N <n>  the reference's option (cmdline.cpp:366-384: integer with K / M / B suffix), here read as a
//                              64-bit count: values above INT_MAX are kept aside for the GPU loop and for the NumPhonons
//                              field of the model-parameter file (model.cpp:148-150, read by vis/seisplot/combine.m:31-33),
//                              while the reference's own parser sees INT_MAX (env R3D_GPU_NUM_PHONONS)
//   --gpu-checkpoint=<path>    checkpoint / exact resume (env R3D_GPU_CHECKPOINT)
//
// r3d_main.cpp parses and removes these before the reference's main() sees the command line; the environment variables
// of round 1 keep working and are overridden by the options.
#ifndef R3D_CLI_HPP_
#define R3D_CLI_HPP_
#include <stdint.h>
#include <string>
#include <vector>

namespace r3d_cli {
struct Options {
  std::vector<int> devices;
  bool have_seed; uint64_t seed;
  bool have_nph; uint64_t nph;
  std::string checkpoint;
  Options() : have_seed(false), seed(0), have_nph(false), nph(0) {}
};
Options & options();            // parsed once (r3d_main.cpp), environment variables as fallback
int first_device();
bool parse_count(const std::string & text, uint64_t & out);     // "125000000", "125M", "10B", "1e10"
}
#endif
