import json, sys
for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "unreadable:", e); print(open(path).read()[-1500:]); continue
    print(path, "VALUE", round(d["value"], 1), "ms/step", round(d["ms_per_step"], 1), "e2e", round(d["e2e"]["value"], 1),
          "gemm TF/s", round(d["roofline"]["achieved"]), "frac", round(d["roofline"]["frac"], 3), "whole TF/s",
          round(d["roofline"]["whole_step_tflops"]), "tile", d["config"]["batch_tile"], d["clocks"])
    for k, v in list(d["kernel_breakdown"].items())[:16]:
        print(f"  {k:14s} {v['ms_per_step']:9.2f} ms  {('%5.0f TF/s' % v['tflops']) if v['tflops'] else ''} {('%5.0f GB/s' % v['gbs']) if v['gbs'] else ''}")
