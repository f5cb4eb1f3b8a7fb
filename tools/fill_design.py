"""Render DESIGN.md from tools/DESIGN.md.in: the @NAME@ fields of its tables are filled from the bench line kept under
profiles/ (the JSON line of `python bench.py --steps 20 --warmup 5` on a B200, see profiles/README.md), so the document and
the committed evidence cannot drift apart.  usage: python tools/fill_design.py [profiles/r02_bench_line.json]"""
import json, os, re, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
line = json.load(open(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_bench_line.json")))
ok = line["other_kernels"]
F = {}


def kernel(prefix, ms, samples_per_s, frac):
    F[prefix + "_MS"] = f"{ms:.3f}"
    F[prefix + "_TS"] = f"{samples_per_s / 1e12:.2f}"
    F[prefix + "_FRAC"] = f"{100 * frac:.1f} %"


kernel("SIMPLE", line["roofline"]["kernel_ms"], line["value"], line["roofline"]["frac"])
F["SIMPLE_APAS"] = f"{line['real_time_apas']:.0f}"
for prefix, key in (("STD", "wibeth_standard_rs"), ("ABS", "wibeth_abs_rs"), ("FIR", "wibeth_fir_iqr_thr5"), ("FIRO", "wibeth_fir_iqr_other_taps"),
                    ("STRESS", "wibeth_simple_stress_thr8"), ("W2S", "wib2_simple"), ("W2F", "wib2_fir_iqr_thr5"), ("W2A", "wib2_abs_rs")):
    kernel(prefix, ok[key]["kernel_ms"], ok[key]["value"], ok[key]["roofline"]["frac"])
e = line["e2e"]
F["E2E_GBS"] = f"{e['h2d_gbs_per_gpu']:.1f}"
F["E2E_GS"] = f"{e['value'] / 1e9:.1f}"
F["E2E_APAS"] = f"{e['real_time_apas']:.2f}"
F["E2E_CORES"] = f"{e['host_cores_busy_per_gpu']:.2f}"
F["E2E_CSAS"] = f"{e['core_seconds_per_apa_second']:.2f}"
F["E2E_FRAC"] = f"{100 * e['ingest_roofline']['frac']:.0f} %"
F["H2D_PEAK"] = f"{e['ingest_roofline']['peak']:.1f}"
F["CLOCK_MHZ"] = str(line["clocks"]["sm_mhz"])
F["CPU_GS"] = f"{line['cpu_baseline']['value'] / 1e9:.1f}"
F["CPU_CORES"] = str(line["cpu_baseline"]["cores"])
F["BATCH_GS"] = f"{line['e2e_batch']['value'] / 1e9:.1f}"
F["BATCH_GBS"] = f"{line['e2e_batch']['h2d_gbs_per_gpu']:.1f}"
F["APA_X"] = f"{line['single_apa']['real_time_multiple']:.1f}"
for n, v in line["link_count_sweep"].items():
    F[f"SWEEP_{n}"] = f"{100 * v['roofline_frac']:.0f} %"
F["SORT_MS"] = f"{line['module_split']['host_sort_ms_per_gpu']:.0f}"
F["DEVSORT_MS"] = f"{line['module_split']['device_sort_ms']:.2f}" if "device_sort_ms" in line["module_split"] else "n/a"
F["SPLIT_X"] = f"{line['module_split']['real_time_multiple_of_the_module']:.1f}"
src = open(os.path.join(ROOT, "tools", "DESIGN.md.in")).read()
missing = sorted(set(re.findall(r"@([A-Z0-9_]+)@", src)) - set(F))
if missing:
    sys.exit(f"no value for {missing}")
open(os.path.join(ROOT, "DESIGN.md"), "w").write(re.sub(r"@([A-Z0-9_]+)@", lambda m: F[m.group(1)], src))
print("DESIGN.md written;", len(F), "fields")
