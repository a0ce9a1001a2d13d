// Host-only invariants of the tile planners (no GPU): every plan tc2_plan / tc3_plan can produce for the shipped layer
// shapes and a sweep of odd ones must fit the hardware (shared memory, TMEM columns, barrier slots) and be internally
// consistent.  Built and run by tests/test_host.py::test_tile_planners_respect_hardware_limits.
#include <atomic>
#include <cstdio>
#include "../../fastvocoder_b200/csrc/fv_tc.cuh"
namespace fv { std::atomic<long long> g_launches{0}; std::atomic<long long> g_tc_launches{0}; }
using namespace fv;

static int failures = 0;
#define CHECK(cond, ...)                                  \
  do {                                                    \
    if (!(cond)) {                                        \
      ++failures;                                         \
      printf("FAIL %s: ", #cond);                         \
      printf(__VA_ARGS__);                                \
      printf("\n");                                       \
    }                                                     \
  } while (0)

static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

static int check_tc2(int B, int Cin, int N, int K, int dil, int L, bool res, int acc, int layout) {
  ConvArgs a{};
  a.B = B; a.Cin = Cin; a.N = N; a.K = K; a.dil = dil; a.Lin = L; a.Lpos = L; a.out_layout = layout;
  a.ph_cout = N; a.ph_lout = L; a.ph_stride = 1;
  a.res = res ? (const float*)16 : nullptr; a.acc_mode = acc;
  TcLayer t; t.eligible = true; t.n_pad = (N + 15) / 16 * 16; t.NT = tc_pick_nt(t.n_pad);
  if (!t.NT || Cin % 16) return 0;
  t.n_tiles = t.n_pad / t.NT;
  Tc2Args p{};
  if (!tc2_plan(a, t, p, 148)) return 0;
  const size_t smem = tc2_smem_bytes(p);
  CHECK(smem <= 227 * 1024, "tc2 Cin=%d N=%d K=%d dil=%d smem=%zu", Cin, N, K, dil, smem);
  CHECK(pow2(p.tmem_cols) && p.tmem_cols >= 32 && p.tmem_cols <= 512, "tmem_cols=%d", p.tmem_cols);
  CHECK(p.acc_stages * p.acc_cols <= p.tmem_cols, "acc %d x %d > %d", p.acc_stages, p.acc_cols, p.tmem_cols);
  CHECK(p.acc_cols == p.m_tiles * p.NT * (p.dual ? 2 : 1), "acc_cols");
  CHECK(p.a_stages >= 1 && p.a_stages <= 2 && p.acc_stages >= 1 && p.acc_stages <= 2, "stages");
  CHECK(p.w_resident || (p.w_stages >= 2 && p.w_stages <= TC_MAX_STAGES), "w_stages=%d", p.w_stages);
  CHECK(p.n_issuers >= 1 && p.n_issuers <= (p.w_resident ? TC2_ISSUE_WARPS : TC2_ISSUE_WARPS - 1), "issuers=%d", p.n_issuers);
  CHECK((p.m_tiles + p.n_issuers - 1) / p.n_issuers <= 4, "M tiles per issuer");
  CHECK(p.rows == p.m_tiles * 128 + (K - 1) * dil, "rows");
  CHECK(p.ck * p.nck == Cin && p.ck % 16 == 0, "chunks");
  CHECK(p.ld_per >= 0 && p.ld_per <= LD_MAX, "ld_per=%d", p.ld_per);
  CHECK(!p.dual || 2 * p.NT <= 256, "dual N");
  CHECK(p.total_tiles == p.tiles_per_batch * B && p.tiles_per_batch * p.m_tiles * 128 >= L, "tiles");
  // upsample (phase-interleaved) layers are epilogue-bound: they must get two accumulator sets whenever one M tile fits twice
  if (layout == OUT_PHASE && !getenv("FV_TC2_UPS_ACC2"))
    CHECK(p.acc_stages == 2 || 2 * p.NT * (p.dual ? 2 : 1) > 512, "upsample plan with one accumulator set: NT=%d mt=%d dual=%d", p.NT,
          p.m_tiles, p.dual);
  return 1;
}

static int check_tc3(int B, int C, int L, int K, int dil) {
  Tc3Args p{};
  if (!tc3_plan(B, C, L, K, dil, p)) return 0;
  const size_t smem = tc3_smem_bytes(p);
  CHECK(smem <= 227 * 1024, "tc3 C=%d K=%d dil=%d smem=%zu", C, K, dil, smem);
  CHECK(pow2(p.tmem_cols) && p.tmem_cols >= 32 && p.tmem_cols <= 512, "tc3 C=%d K=%d tmem_cols=%d", C, K, p.tmem_cols);
  const int sets = p.pp ? 2 : p.acc1_stages + p.acc2_stages;
  CHECK(sets * p.acc_cols <= p.tmem_cols, "tc3 C=%d K=%d sets %d x %d > %d", C, K, sets, p.acc_cols, p.tmem_cols);
  CHECK(p.acc_cols == p.m_tiles * 2 * C, "acc_cols");
  CHECK(p.m_out == 128 * p.m_tiles - (K - 1) && p.m_out > 0, "m_out=%d", p.m_out);
  CHECK(p.x_rows == 128 * p.m_tiles + (K - 1) * dil && p.h_rows_alloc == 128 * p.m_tiles + (K - 1), "rows");
  CHECK(p.a1_stages >= 1 && p.a1_stages <= 2 && p.acc1_stages >= 1 && p.acc1_stages <= 2 && p.acc2_stages >= 1 &&
            p.acc2_stages <= 2, "stages");
  CHECK(!p.pp || (p.a1_stages == 2 && p.acc1_stages == 2 && p.acc2_stages == 2 && p.x_rows >= p.h_rows_alloc), "pp invariants");
  CHECK(p.w_resident || (p.w_stages >= 3 && p.w_stages <= 4 && p.stage_bytes == p.ksteps * C * 64 && p.stage_bytes % 16 == 0),
        "ring: w_stages=%d stage_bytes=%d", p.w_stages, p.stage_bytes);
  CHECK(p.n_issuers >= 1 && p.n_issuers <= (p.w_resident ? TC2_ISSUE_WARPS : TC2_ISSUE_WARPS - 1), "issuers=%d", p.n_issuers);
  CHECK(p.ld_per >= 0 && p.ld_per <= LD_MAX, "ld_per");
  CHECK((long long)p.tiles_per_batch * p.m_out >= L, "coverage");
  return 1;
}

// fused ResidualStack plans (tc3_plan_stack): ping-pong tiles, resident images of different sizes, raw planes behind the x planes
static int check_stack(int B, int C, int L, int K, int dil) {
  Tc3Args p{};
  if (!tc3_plan_stack(B, C, L, K, dil, p)) return 0;
  const size_t smem = tc3_smem_bytes(p);
  CHECK(smem <= 227 * 1024, "stack C=%d K=%d dil=%d smem=%zu", C, K, dil, smem);
  CHECK(pow2(p.tmem_cols) && p.tmem_cols >= 32 && p.tmem_cols <= 512, "stack C=%d tmem_cols=%d", C, p.tmem_cols);
  CHECK(2 * p.acc_cols <= p.tmem_cols && p.acc_cols == p.m_tiles * 2 * C, "stack acc %d cols, tmem %d", p.acc_cols, p.tmem_cols);
  CHECK(p.stack == 1 && p.pp == 1 && p.w_resident == 1 && p.a1_stages == 2 && p.acc1_stages == 2 && p.acc2_stages == 2, "stack mode");
  CHECK(p.k2 == 1 && p.ksteps2 == 2 * p.ksteps && p.kblocks2 == 2 * p.ksteps && p.kblocks == K * p.ksteps, "stack images");
  CHECK(p.m_out == 128 * p.m_tiles && p.x_rows == p.m_out + (K - 1) * dil && p.x_rows_alloc >= p.x_rows, "stack rows");
  CHECK(p.n_issuers >= 1 && p.n_issuers <= TC2_ISSUE_WARPS && p.m_tiles <= p.n_issuers * 4, "stack issuers=%d", p.n_issuers);
  CHECK(p.ld_per >= 1 && p.ld_per <= LD_MAX && p.ld_rounds >= 1 && 256 * p.ld_per * p.ld_rounds >= (C / 8) * p.x_rows, "stack loader plan");
  CHECK((long long)p.tiles_per_batch * p.m_out >= L && p.total_tiles == p.tiles_per_batch * B, "stack coverage");
  return 1;
}

int main() {
  int planned = 0;
  // every conv-like shape of the shipped configs (channels, taps, dilations), plus odd lengths and batch sizes
  const int Ls[] = {1, 7, 131, 1000, 8000, 40000, 120001, 240000};
  const int Bs[] = {1, 3, 32, 65};
  for (int L : Ls)
    for (int B : Bs) {
      for (int C : {16, 32, 48, 64, 128, 256, 512})
        for (int K : {1, 3, 7, 11})
          for (int d : {1, 3, 5, 9}) {
            for (int res = 0; res < 2; ++res)
              for (int acc : {ACC_STORE, ACC_ADD, ACC_ADD_DIV, ACC_STORE_SCALE, ACC_RED_SCALE})
                planned += check_tc2(B, C, C, K, d, L, res != 0, acc, OUT_BCL);
            if (C <= 64 && K > 1) planned += check_tc3(B, C, L, K, d);
            if (C <= 64 && K == 3) planned += check_stack(B, C, L, K, d);
          }
      planned += check_tc2(B, 80, 256, 7, 1, L, false, ACC_STORE, OUT_BCL);      // conv_pre
      planned += check_tc2(B, 80, 512, 7, 1, L, false, ACC_STORE, OUT_BCL);
      for (int s : {2, 3, 4, 5, 6, 8, 10})                                       // polyphase ConvTranspose / UpsampleLayer
        for (int C : {32, 64, 128, 256, 512})
          planned += check_tc2(B, C, s * (C / 2), 2, 1, L, false, ACC_STORE, OUT_PHASE);
      planned += check_tc2(B, 256, 15, 2, 1, L, false, ACC_STORE, OUT_BLC);      // basis linear + overlap-add
      planned += check_tc2(B, 512, 256, 1, 1, L, false, ACC_STORE, OUT_BCL);     // ResidualStack pair
      planned += check_tc2(B, 16, 1, 7, 1, L, false, ACC_STORE, OUT_BCL);        // conv_post
      planned += check_tc2(B, 64, 4, 7, 1, L, false, ACC_STORE, OUT_BCL);
    }
  printf("plans checked: %d, failures: %d\n", planned, failures);
  return failures ? 1 : 0;
}
