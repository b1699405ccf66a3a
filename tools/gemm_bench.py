"""Times the contraction kernel on the workload's main shapes (CUDA events on the launching stream)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from shap_transformer_asr_b200 import debug_gemm

SHAPES = {  # name: (M, N, K, act, out_fp32)   C2 at batch tile 64
    "ffn1": (15936, 3072, 768, 1, 0), "ffn2": (15936, 768, 3072, 0, 1), "qkv": (15936, 2304, 768, 0, 0),
    "out_proj": (15936, 768, 768, 0, 1), "conv1_like": (64 * 7999, 512, 1536, 1, 0), "conv6_like": (15936, 512, 1024, 1, 0),
}
only = sys.argv[1:] or list(SHAPES)
res = {}
for name in only:
    M, N, K, act, f32 = SHAPES[name]
    a = torch.randn(M, K, device="cuda").bfloat16()
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).bfloat16()
    bias = torch.randn(N, device="cuda")
    for _ in range(3):
        debug_gemm(a, w, bias, act=act, out_fp32=bool(f32))
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 10
    e0.record()
    for _ in range(n):
        debug_gemm(a, w, bias, act=act, out_fp32=bool(f32))
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res[name] = dict(ms=ms, tflops=2.0 * M * N * K / ms / 1e9)
    print(name, res[name], flush=True)
