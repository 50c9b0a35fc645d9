/* mixed_999.cu -- the M = 999 = 9 * 3 * 37 instances of the mixed-radix streaming kernel (mixed.cuh). */
#include "mixed_impl.cuh"

namespace cfb {
typedef MixCfg<9, 3, 37> C999;
bool mix_launch_999(int kind, int dir, long long npairs, double *x, const double *trig) {
  return mix_launch_cfg<C999>(kind, dir, npairs, x, trig);
}
void mix_release_999() { MixTables<C999>::release(); }
}  // namespace cfb
