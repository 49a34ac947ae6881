// Event timeline of one attention-v4 CTA (tiles 40..55): when the MMA thread issues S_t(n) / PV_t(n) and when each
// softmax warpgroup passes its hand-off points, in SM clocks.  Builds the kernel sources with DOVE_ATTN_TRACE.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr -Iinclude \
//        -o profiles/bin/trace_attn profiles/trace_attn.cu && profiles/bin/trace_attn <variant>
#define DOVE_ATTN_TRACE 1
#include "../dove_b200/csrc/common.cu"
#include "../dove_b200/csrc/attn.cu"
#include <algorithm>
#include <vector>

int main(int argc, char** argv) {
  const int variant = argc > 1 ? atoi(argv[1]) : 7;
  const int rows = 19426, heads = 48;
  if (dove_init(0)) { printf("%s\n", dove_last_error()); return 1; }
  std::vector<uint16_t> h(static_cast<size_t>(rows) * 3 * heads * 64);
  uint32_t s = 12345;
  for (auto& x : h) {   // bf16 in (-2, 2), crude
    s = s * 1664525u + 1013904223u;
    const float f = (static_cast<int>(s >> 8) % 4001 - 2000) * 1e-3f;
    uint32_t u; memcpy(&u, &f, 4);
    x = static_cast<uint16_t>(u >> 16);
  }
  void *qkv, *out;
  cudaMalloc(&qkv, h.size() * 2);
  cudaMalloc(&out, static_cast<size_t>(rows) * heads * 64 * 2);
  cudaMemcpy(qkv, h.data(), h.size() * 2, cudaMemcpyHostToDevice);
  dove_set_option("attn_variant", variant);
  for (int it = 0; it < 3; ++it) {
    int zero[4] = {0, 0, 0, 0};
    cudaMemcpyToSymbol(dove::g_attn_trace_n, zero, sizeof(zero));
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    if (dove_attention_bf16(qkv, out, rows, heads, 0.125f, nullptr)) { printf("%s\n", dove_last_error()); return 1; }
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    printf("# variant %d iteration %d: %.3f ms\n", variant, it, ms);
  }
  static unsigned long long tr[4][1024];
  int n[4];
  cudaMemcpyFromSymbol(tr, dove::g_attn_trace, sizeof(tr));
  cudaMemcpyFromSymbol(n, dove::g_attn_trace_n, sizeof(n));
  struct Ev { unsigned long long clk; int actor, ev, tile; };
  std::vector<Ev> evs;
  for (int a = 0; a < 4; ++a)
    for (int k = 0; k < std::min(n[a], 1024); ++k)
      evs.push_back({tr[a][k] & 0xffffffffffull, a, static_cast<int>((tr[a][k] >> 40) & 0xff), static_cast<int>(tr[a][k] >> 48)});
  std::sort(evs.begin(), evs.end(), [](const Ev& x, const Ev& y) { return x.clk < y.clk; });
  const char* names[16] = {"", "issue S0", "issue S1", "issue PV0", "issue PV1", "", "", "", "", "wait s_full", "got s_full",
                           "wait pv_done", "got pv_done", "s_free", "p_full", ""};
  const unsigned long long t0 = evs.empty() ? 0 : evs[0].clk;
  for (const Ev& e : evs)
    printf("%8llu  %-6s %-13s tile %d\n", e.clk - t0, e.actor == 0 ? "MMA0" : e.actor == 1 ? "MMA1" : e.actor == 2 ? "WG0" : "WG1", names[e.ev & 15], e.tile);
  return 0;
}
