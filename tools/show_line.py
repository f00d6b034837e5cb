"""one-line summary of a bench.py JSON line (tuning runs)"""
import json
import sys

for path in sys.argv[1:]:
    for ln in open(path):
        if not ln.startswith("{"):
            continue
        d = json.loads(ln)
        r = d.get("roofline", {})
        e = d.get("e2e") or {}
        p = d.get("parity_vs_n1") or {}
        print(path.split("/")[-1], "value %.1f e2e %.1f (%.0f ms) set_graph %.3f frac %s phases_us %s step_ms %s parity %s clocks %s %s" % (
            d["value"], e.get("value", 0), e.get("ms_per_step", 0), d.get("set_graph_s", 0),
            ("%.3f" % r["frac"]) if r.get("frac") else None,
            [round(x, 1) for x in r.get("pcg_phase_us_per_iteration", [])],
            {k: round(v / max(1, d["steps"]), 1) for k, v in d.get("phases_ms", {}).items()},
            p.get("ok"), d["clocks"].get("sm_mhz"), d["clocks"].get("reasons")))
