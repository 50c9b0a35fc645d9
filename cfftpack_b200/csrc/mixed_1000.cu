/* mixed_1000.cu -- the M = 1000 = 10 * 10 * 10 instances of the mixed-radix streaming kernel (mixed.cuh). */
#include "mixed_impl.cuh"

namespace cfb {
typedef MixCfg<10, 10, 10> C1000;
bool mix_launch_1000(int kind, int dir, long long npairs, double *x, const double *trig) {
  return mix_launch_cfg<C1000>(kind, dir, npairs, x, trig);
}
void mix_release_1000() { MixTables<C1000>::release(); }
}  // namespace cfb
