/* mixed.cu -- dispatch of the mixed-radix streaming kernel by underlying real-transform length. */
#include "mixed.cuh"

namespace cfb {

#define CFB_MIX_DECL(M)                                                                        \
  bool mix_launch_##M(int kind, int dir, long long npairs, double *x, const double *trig);    \
  void mix_release_##M();
CFB_MIX_DECL(1001)
CFB_MIX_DECL(1000)
CFB_MIX_DECL(999)
CFB_MIX_DECL(1002)

static int underlying_length(int kind, int n) { return kind == K_COST ? n - 1 : kind == K_SINT ? n + 1 : n; }

bool mix_supported(int kind, int n) {
  if (kind != K_RFFT && kind != K_COSQ && kind != K_SINT && kind != K_COST) return false;
  const int M = underlying_length(kind, n);
  return M == 999 || M == 1000 || M == 1001 || M == 1002;
}

bool mix_launch(int kind, int n, long long npairs, int dir, double *x, const double *trig) {
  if (npairs <= 0) return true;
  switch (mix_supported(kind, n) ? underlying_length(kind, n) : 0) {
    case 1001: return mix_launch_1001(kind, dir, npairs, x, trig);
    case 1000: return mix_launch_1000(kind, dir, npairs, x, trig);
    case 999: return mix_launch_999(kind, dir, npairs, x, trig);
    case 1002: return mix_launch_1002(kind, dir, npairs, x, trig);
    default: break;
  }
  set_error("mix_launch: unsupported (kind %d, n %d)", kind, n);
  return false;
}

void mix_release_tables() {
  mix_release_1001();
  mix_release_1000();
  mix_release_999();
  mix_release_1002();
}

}  // namespace cfb
