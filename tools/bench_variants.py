"""Developer tool: time the kernel variants selected by CFB200_POW2_VARIANT (one process per variant)."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
variants = sys.argv[1].split(",") if len(sys.argv) > 1 else ["0", "1", "2", "3", "4", "5"]
extra = sys.argv[2:] 
for v in variants:
    env = dict(os.environ, CFB200_POW2_VARIANT=v)
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "10", "--warmup", "3", "--no-cpu-baseline",
                          "--no-e2e"] + extra, env=env, capture_output=True, text=True)
    try:
        j = json.loads(out.stdout.strip().splitlines()[-1])
        print(f"variant {v}: {j['ms_per_step']:.3f} ms/step  {j['value']:.0f} GB/s  frac_measured={j['roofline']['frac']:.3f} frac_8TB={j['frac_of_8TBps']:.3f}", flush=True)
    except Exception as e:
        print("variant", v, "failed", out.stdout[-500:], out.stderr[-1500:])
