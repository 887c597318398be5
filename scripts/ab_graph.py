"""A/B timing of the graphed training step under different environment settings, interleaved in ONE process
launch sequence so that box-to-box variation cancels:  python scripts/ab_graph.py "A=1" "B=1 C=2" ...
Each setting is run `rounds` times (default 3) in round-robin order; prints the per-setting median."""
import os, subprocess, sys, statistics
settings = sys.argv[1:] or [""]
rounds = int(os.environ.get("AB_ROUNDS", "3"))
here = os.path.dirname(os.path.abspath(__file__))
res = {s: [] for s in settings}
for r in range(rounds):
    for s in settings:
        env = dict(os.environ)
        for kv in s.split():
            k, v = kv.split("=", 1)
            env[k] = v
        out = subprocess.run([sys.executable, os.path.join(here, "time_graph.py"), "graph-only"], env=env,
                             capture_output=True, text=True, timeout=600).stdout
        for ln in out.splitlines():
            if ln.startswith("B=32") and " graph:" in ln:
                res[s].append(float(ln.split("graph:")[1].split("ms/step")[0]))
for s in settings:
    v = res[s]
    print(f"[{s or 'default'}] median {statistics.median(v):.3f} ms/step  runs {['%.3f' % x for x in v]}", flush=True)
