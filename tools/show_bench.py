"""Pretty-prints the JSON line bench.py emits (reads stdin)."""
import json, sys
txt = sys.stdin.read().strip().splitlines()
line = next((l for l in reversed(txt) if l.startswith("{")), None)
if line is None:
    print("\n".join(txt[-30:]))
    sys.exit(1)
d = json.loads(line)
for k in ["value", "ms_per_step", "wall_ms_per_step", "gpu_launches"]:
    print(k, d.get(k))
print("e2e", d.get("e2e"))
print("roofline", d.get("roofline"))
print("cpu_baseline", d.get("cpu_baseline"))
for k in ["stage_ms_per_step", "host_ms_per_step", "op_ms_per_step"]:
    print(k)
    for a, b in sorted((d.get(k) or {}).items(), key=lambda kv: -kv[1]):
        print("    %-36s %9.3f" % (a, b))
