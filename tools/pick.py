"""Print the headline fields of one bench.py JSON line read from stdin (used for A/B runs inside one gpurun call)."""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
d = json.loads([l for l in sys.stdin.read().splitlines() if l.startswith("{")][-1])   # torchrun prints a banner first
r = d.get("roofline", {})
print(tag, f"value={d['value']:.0f} ms={d['ms_per_step']:.3f} e2e={d['e2e']['value']:.0f} loss={d['config'].get('loss')}",
      f"bwd_ms={r.get('launch_ms')} fwd_ms={r.get('fwd_sweep', {}).get('launch_ms')} clocks={d.get('clocks')}")
