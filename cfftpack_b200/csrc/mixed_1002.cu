/* mixed_1002.cu -- the M = 1002 = 6 * 167 instances of the mixed-radix streaming kernel (mixed.cuh). */
#include "mixed_impl.cuh"

namespace cfb {
typedef MixCfg<6, 1, 167> C1002;
bool mix_launch_1002(int kind, int dir, long long npairs, double *x, const double *trig) {
  return mix_launch_cfg<C1002>(kind, dir, npairs, x, trig);
}
void mix_release_1002() { MixTables<C1002>::release(); }
}  // namespace cfb
