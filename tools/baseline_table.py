"""Markdown rows for BASELINE.md from the committed bench lines under profiles/.   python tools/baseline_table.py profiles/bench_c2_r02k.json ..."""
import json
import sys


def last_json(path):
    lines = [l for l in open(path).read().strip().splitlines() if l.startswith("{")]
    return json.loads(lines[-1])


for p in sys.argv[1:]:
    o = last_json(p)
    r, cfg, e = o["roofline"], o["config"], o["e2e"]
    s = o.get("stage_ms_per_step", {})
    pc = o.get("parity") or o.get("parity_check") or {}
    print(f"| {cfg['workload'].split(':')[0]} x{o['n_gpus']} | {cfg.get('nprobe')} | {cfg.get('recall_at_10')} | {o['value'] / 1e6:.3f} M | {e['value'] / 1e6:.3f} M | "
          f"{o['ms_per_step']:.3f} | {r['gpairs_per_s']} | {r['achieved']:.0f} | {r['frac']:.2f} | "
          f"{'all-true' if pc.get('dist_bit_identical') and pc.get('ids_identical_up_to_ties') and pc.get('precise_equal') else pc} | `{p}` |")
    if s:
        print("    stages: " + ", ".join(f"{k} {v}" for k, v in s.items()))
