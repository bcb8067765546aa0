"""GPU debug helper: runs conv cases on the tcgen05 engine against torch CPU fp32 and prints error
statistics.  Each case runs in its own subprocess under a timeout (a protocol bug traps or times
out instead of wedging the session).  Usage: python tools/tc_debug.py [case-filter] """
import math, os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = {
    "c64_k3": (64, 64, 3, 1, 0, 0, 16, 32),
    "c64_k3_big": (64, 64, 3, 1, 0, 1, 70, 100),
    "c128_k3": (128, 128, 3, 1, 0, 2, 34, 60),
    "c32_k7": (32, 64, 7, 1, 0, 1, 34, 60),
    "c64_k7": (64, 32, 7, 1, 0, 1, 34, 60),
    "c8_k7": (8, 32, 7, 1, 0, 1, 24, 40),
    "c16_k7_o2": (16, 2, 7, 1, 0, 0, 20, 20),
    "c128_k3_s2": (128, 128, 3, 2, 0, 2, 32, 64),
    "c2_k3_s2": (2, 128, 3, 2, 0, 2, 32, 48),
    "c128_k3_t2": (128, 128, 3, 2, 1, 2, 17, 30),
    "c64_k5_s2": (64, 96, 5, 2, 0, 0, 32, 64),
    "c96_k5_t2": (96, 64, 5, 2, 1, 0, 8, 16),
    "c64_k3_t1": (64, 96, 3, 1, 1, 3, 8, 14),
    "c64_o3": (64, 3, 3, 1, 0, 0, 16, 16),
}

def run_case(name):
    import torch, torch.nn.functional as F
    from fastvideocodec_b200 import ops
    cin, cout, k, stride, tr, act, H, W = CASES[name]
    g = torch.Generator().manual_seed(1)
    x = torch.randn((1, cin, H, W), generator=g)
    wshape = (cin, cout, k, k) if tr else (cout, cin, k, k)
    w = torch.randn(wshape, generator=g) / math.sqrt(cin * k * k)
    b = torch.randn((cout,), generator=g) * 0.1
    if tr:
        want = F.conv_transpose2d(x, w, b, stride=stride, padding=k // 2, output_padding=stride - 1)
    else:
        want = F.conv2d(x, w, b, stride=stride, padding=k // 2)
    want = {0: lambda t: t, 1: torch.relu, 2: lambda t: F.leaky_relu(t, 0.1), 3: torch.exp}[act](want)
    dev = torch.device("cuda")
    fn = ops.conv_transpose2d if tr else ops.conv2d
    got = fn(x.to(dev), w.to(dev), b.to(dev), stride, act, ops.IMPL_TC).cpu()
    simt = fn(x.to(dev), w.to(dev), b.to(dev), stride, act, ops.IMPL_SIMT).cpu()
    err = (got - want).abs()
    rel = err.max().item() / max(1.0, want.abs().max().item())
    bad = (err > 1e-3 * max(1.0, want.abs().max().item())).float().mean().item()
    print(json.dumps({"case": name, "bo": os.environ.get("FVC_TC_BO_MODE", "0"), "pw": os.environ.get("FVC_TC_PW_ALIGN", "8"),
                      "max_rel_err": rel, "frac_bad": bad, "simt_rel": (simt - want).abs().max().item() / max(1.0, want.abs().max().item()),
                      "mean_abs_err": err.mean().item(), "nan": bool(torch.isnan(got).any())}))
    if rel > 1e-4 and os.environ.get("FVC_TC_DUMP"):
        e = err[0].max(0).values  # [Ho,Wo] max over channels
        rows = (e > 1e-3).nonzero()
        print("bad pixels (first 20):", rows[:20].tolist(), "per-channel bad:", (err[0] > 1e-3).flatten(1).any(1).nonzero().flatten()[:16].tolist())

if __name__ == "__main__":
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        run_case(sys.argv[2]); sys.exit(0)
    flt = sys.argv[1] if len(sys.argv) > 1 else ""
    for name in CASES:
        if flt and flt not in name: continue
        for bo in os.environ.get("TC_BOS", "0").split(","):
            env = dict(os.environ, FVC_TC_BO_MODE=bo)
            try:
                r = subprocess.run([sys.executable, __file__, "--one", name], env=env, capture_output=True, text=True, timeout=120)
                out = (r.stdout.strip().splitlines() or ["<no output>"])
                print("\n".join(out[-3:]) if r.returncode == 0 else "FAIL %s bo=%s rc=%d: %s | %s" % (name, bo, r.returncode, out[-1][:300], r.stderr.strip()[-400:]))
            except subprocess.TimeoutExpired:
                print("TIMEOUT %s bo=%s" % (name, bo))
            sys.stdout.flush()
