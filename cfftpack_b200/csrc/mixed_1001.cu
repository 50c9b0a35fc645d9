/* mixed_1001.cu -- the M = 1001 = 13 * 11 * 7 instances of the mixed-radix streaming kernel (mixed.cuh). */
#include "mixed_impl.cuh"

namespace cfb {
typedef MixCfg<13, 11, 7> C1001;
bool mix_launch_1001(int kind, int dir, long long npairs, double *x, const double *trig) {
  return mix_launch_cfg<C1001>(kind, dir, npairs, x, trig);
}
void mix_release_1001() { MixTables<C1001>::release(); }
}  // namespace cfb
