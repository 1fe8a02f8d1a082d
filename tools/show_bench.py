"""Print the headline fields of bench.py JSON lines: python tools/show_bench.py file.json ..."""
import json
import sys

for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:  # noqa: BLE001
        print(path, "unreadable:", e)
        continue
    r = d.get("roofline", {}) or {}
    print(f"{path}: n_gpus={d.get('n_gpus')} ms/step={d['ms_per_step']:.2f} value={d['value']:.4g} "
          f"e2e_ms={d['e2e'].get('ms_per_step', float('nan')):.2f} e2e={d['e2e']['value']:.4g} "
          f"kernel_ms={r.get('kernel_ms')} TOPS={r.get('achieved')} frac={r.get('frac')} "
          f"nominal_frac={r.get('frac_of_nominal')} of_int8_nominal={r.get('frac_of_nominal_int8')} operand={d['config'].get('operand')}")
    print("   clocks:", d.get("clocks"), "phases:", d.get("phases_ms"), "cpu:", d.get("cpu_baseline"))
    print("   host phases:", d["e2e"].get("host_phases_ms"), "parity:", d.get("parity_check"),
          "parallelism:", d["config"].get("parallelism"))
