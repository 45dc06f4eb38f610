"""Key fields of a bench.py JSON line: python scripts/bench_digest.py gpurun_out/bench.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
r, e = d.get("roofline", {}), d.get("extra", {})
print(f"n_gpus {d.get('n_gpus')} steps {d.get('steps')} warmup {d.get('warmup')}: value {d['value']:.4g} {d['unit']} ({d['ms_per_step']:.4f} ms/step), "
      f"e2e {d['e2e']['value']:.4g}, gpu_launches {d.get('gpu_launches')}, clocks {d.get('clocks', {}).get('sm_mhz')} MHz {d.get('clocks', {}).get('reasons')}")
print(f"  roofline: bound {r.get('bound')}, frac {r.get('frac'):.4f} (alone {r.get('frac_alone', 0):.4f}), kernels_us {r.get('kernels_us')}")
cb = d.get("cpu_baseline") or {}
print(f"  cpu_baseline: {cb.get('value')} on {cb.get('cores')} cores ({cb.get('kind')}); python reference {((cb.get('python_reference') or {}).get('value'))}")
for k, v in e.items():
    if isinstance(v, dict):
        keep = {kk: (round(vv, 4) if isinstance(vv, float) else vv) for kk, vv in v.items() if isinstance(vv, (int, float)) or kk.endswith("_all_this_rank")}
        print(f"  {k}: {json.dumps(keep)[:400]}")
