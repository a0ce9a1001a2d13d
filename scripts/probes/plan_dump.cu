// Host-only: print the tile plans tc2_plan / tc3_plan choose for the BASELINE layer shapes (no GPU needed).
#include <atomic>
#include <cstdio>
#include "../../fastvocoder_b200/csrc/fv_tc.cuh"
namespace fv { std::atomic<long long> g_launches{0}; std::atomic<long long> g_tc_launches{0}; }
using namespace fv;
static void show(const char* name, int B, int Cin, int N, int K, int dil, int L, bool res, int acc, int layout = OUT_BCL) {
  ConvArgs a{};
  a.B = B; a.Cin = Cin; a.N = N; a.K = K; a.dil = dil; a.Lin = L; a.Lpos = L; a.out_layout = layout;
  a.res = res ? (const float*)16 : nullptr; a.acc_mode = acc;
  TcLayer t; t.eligible = true; t.n_pad = (N + 15) / 16 * 16; t.NT = tc_pick_nt(t.n_pad); t.n_tiles = t.n_pad / t.NT;
  Tc2Args p{};
  if (!tc2_plan(a, t, p, 148)) { printf("%-28s no plan\n", name); return; }
  printf("%-28s NT=%d mt=%d ck=%d a_st=%d acc_st=%d res=%d w_st=%d kbps=%d iss=%d dual=%d tiles=%d smem=%zu\n", name, p.NT,
         p.m_tiles, p.ck, p.a_stages, p.acc_stages, p.w_resident, p.w_stages, p.kb_per_stage, p.n_issuers, p.dual,
         p.total_tiles, tc2_smem_bytes(p));
}
int main() {
  for (int K : {3, 7, 11}) {
    char n[64];
    snprintf(n, 64, "C128 k%d conv1", K); show(n, 32, 128, 128, K, 3, 8000, false, ACC_STORE);
    snprintf(n, 64, "C128 k%d conv2+res", K); show(n, 32, 128, 128, K, 1, 8000, true, ACC_STORE);
    snprintf(n, 64, "C128 k%d conv2+res+acc", K); show(n, 32, 128, 128, K, 1, 8000, true, ACC_ADD);
  }
  for (int K : {7, 11}) {
    char n[64];
    snprintf(n, 64, "C64 k%d conv1", K); show(n, 32, 64, 64, K, 3, 40000, false, ACC_STORE);
    snprintf(n, 64, "C64 k%d conv2+res", K); show(n, 32, 64, 64, K, 1, 40000, true, ACC_STORE);
    snprintf(n, 64, "C64 k%d conv2+res+acc", K); show(n, 32, 64, 64, K, 1, 40000, true, ACC_ADD);
  }
  show("ups0 256->8x128 k2", 32, 256, 1024, 2, 1, 1001, false, ACC_STORE, OUT_PHASE);
  show("ups1 128->5x64 k2", 32, 128, 320, 2, 1, 8001, false, ACC_STORE, OUT_PHASE);
  show("ups2 64->3x32", 32, 64, 96, 2, 1, 40001, false, ACC_STORE, OUT_PHASE);
  show("ups3 32->2x16", 32, 32, 32, 2, 1, 120001, false, ACC_STORE, OUT_PHASE);
  show("conv_post 16->1 k7", 32, 16, 1, 7, 1, 240000, false, ACC_STORE);
  show("basis C256 k3 d9 @16T", 65, 256, 256, 3, 9, 16000, false, ACC_STORE);
  show("basis pair 512->256", 65, 512, 256, 1, 1, 16000, false, ACC_STORE);
  show("basis convT 256->4x256", 65, 256, 1024, 2, 1, 4001, false, ACC_STORE, OUT_PHASE);
  show("basis lin 256->15 k2", 65, 256, 15, 2, 1, 16001, false, ACC_STORE, OUT_BLC);
  return 0;
}
