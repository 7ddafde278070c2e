"""Builds tools/_liblerax_prof.so: the library with per-step cycle counters compiled into the fused PPO gradient
kernel (-DLRX_FUSED_PROFILE; never part of the product build).  Usage on the GPU box:
    cp tools/_liblerax_prof.so lerax_b200/_C/liblerax_b200.so && python bench.py --steps 1 --warmup 1 --epochs 1 ..."""
import os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = os.path.join(ROOT, "lerax_b200", "csrc")
s = open(os.path.join(CS, "lrx_update_fused.cu")).read()
s = s.replace('''  auto step_impl = [&](auto&& crit, auto&& bg, bool track_bg, bool after_loads = false) {
    if (worker) {''', '''  long long pf_t[20] = {0};
  long long pf_last = clock64();
  int pf_i = 0;
  auto step_impl = [&](auto&& crit, auto&& bg, bool track_bg, bool after_loads = false) {
    {
      const long long now = clock64();
      pf_t[pf_i] += now - pf_last;
      pf_last = now;
      pf_i = pf_i + 1;
    }
    if (worker) {''')
s = s.replace('''  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x,''', '''  extern __shared__ __align__(1024) uint8_t smem[];
  const long long pf_k0 = clock64();
  const int tid = threadIdx.x,''')
s = s.replace('''  int tile_parity = 0;''', '''  const long long pf_k1 = clock64();
  int tile_parity = 0;''')
s = s.replace('''  if (warp == 0) umma::tmem_dealloc(tb, TMEM_COLS);
}''', '''  if (warp == 0) umma::tmem_dealloc(tb, TMEM_COLS);
  if (blockIdx.x == 0 && tid == 0 && a.prof) { a.prof[15] = pf_k1 - pf_k0; a.prof[16] = pf_k2 - pf_k1; a.prof[17] = clock64() - pf_k2; }
}''')
s = s.replace('''  // ------------------------------------------------------------------ flush the gradient partial''', '''  const long long pf_k2 = clock64();
  // ------------------------------------------------------------------ flush the gradient partial''')
s = s.replace('''  // ------------------------------------------------------------------ flush the gradient partial''', '''  if (blockIdx.x == 0 && (tid == 128 || tid == 0 || tid == 512) && a.prof) {
    long long* o = a.prof + (tid == 0 ? 0 : tid == 128 ? 20 : 40);
    for (int i = 0; i < 20; ++i) o[i] = pf_t[i];
  }
  // ------------------------------------------------------------------ flush the gradient partial''')
s = s.replace('''  int obs_dim, n_out, is_box;
  tc::LayerOffsets lo;
};''', '''  int obs_dim, n_out, is_box;
  tc::LayerOffsets lo;
  long long* prof;
};''')
s = s.replace('''  switch (pol->n_out) {''', '''  static long long* prof_dev = nullptr;
  static int prof_left = 2;
  if (!prof_dev) cudaMalloc(&prof_dev, 60 * sizeof(long long));
  a.prof = prof_dev;
  switch (pol->n_out) {''')
s = s.replace('''  LRX_LAUNCH_CHECK();
  lrx_launch_reduce_partials(''', '''  LRX_LAUNCH_CHECK();
  if (prof_left > 0 && a.n_tiles > 1000) {
    long long hp[60];
    cudaStreamSynchronize(stream);
    cudaMemcpy(hp, prof_dev, sizeof(hp), cudaMemcpyDeviceToHost);
    const int tiles0 = (a.n_tiles + grid - 1) / grid;
    const char* names[3] = {"warp0(g0)", "warp4(g1)", "issuer"};
    for (int w = 0; w < 3; ++w) {
      fprintf(stderr, "[fused prof] %s cycles/tile before step k:", names[w]);
      long long tot = 0;
      for (int i = 0; i < 14; ++i) { fprintf(stderr, " %lld", hp[20 * w + i] / tiles0); tot += hp[20 * w + i] / tiles0; }
      fprintf(stderr, "  | total %lld\\n", tot);
    }
    fprintf(stderr, "[fused prof] block 0: prologue %lld, loop %lld, tail %lld cycles; tiles %d; prologue parts: weights %lld, fp32 %lld, alloc+barrier %lld\\n", hp[15], hp[16], hp[17], tiles0, hp[18], hp[19], hp[14]);
    --prof_left;
  }
  lrx_launch_reduce_partials(''', 1)
# marks inside the tile-start section (slots 12, 13: cycles since the last step entry)
s = s.replace('''    tile_parity ^= 1;
''', '''    tile_parity ^= 1;
    pf_i = 0;
    pf_t[12] += clock64() - pf_last;
''')
s = s.replace('''      mk_h1 = 0;
''', '''      pf_t[13] += clock64() - pf_last;
      mk_h1 = 0;
''')
# prologue marks (block 0, thread 0): after the weight staging, the fp32 block, the TMEM allocation barrier
s = s.replace('''  tc::stage_f32(a.params, a.lo, n_out, a.obs_dim, fp, tid, THREADS);
''', '''  const long long pm0 = clock64();
  tc::stage_f32(a.params, a.lo, n_out, a.obs_dim, fp, tid, THREADS);
  const long long pm1 = clock64();
''')
s = s.replace('''  const uint32_t tb = *tslot;
''', '''  const long long pm2 = clock64();
  const uint32_t tb = *tslot;
''')
s = s.replace('''a.prof[17] = clock64() - pf_k2; }''', '''a.prof[17] = clock64() - pf_k2; a.prof[18] = pm0 - pf_k0; a.prof[19] = pm1 - pm0; a.prof[14] = pm2 - pm1; }''')
s = s.replace('''cycles; tiles %d\\\\n", hp[15], hp[16], hp[17], tiles0);''', '''cycles; tiles %d; prologue parts: weights %lld, fp32 %lld, alloc+barrier %lld\\\\n", hp[15], hp[16], hp[17], tiles0, hp[18], hp[19], hp[14]);''')
tmp = "/tmp/lrx_update_fused_prof.cu"
open(tmp, "w").write(s)
nv = "/usr/local/cuda/bin/nvcc"
flags = ["-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-I", CS]
subprocess.check_call([nv] + flags + ["-c", tmp, "-o", "/tmp/lrx_update_fused_prof.o"])
objs = [os.path.join(CS, "build", f) for f in os.listdir(os.path.join(CS, "build")) if f.endswith(".o") and f != "lrx_update_fused.o"]
subprocess.check_call([nv, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", os.path.join(ROOT, "tools", "_liblerax_prof.so"), "/tmp/lrx_update_fused_prof.o"] + objs + ["-lcudart"])
print("built tools/_liblerax_prof.so")
