"""Turn the ncu captures in gpurun_out/ into the tracked summaries under profiles/ (needs the ncu CLI, no GPU).
usage: python tools/make_profiles.py <round tag, e.g. r01>"""
import csv, json, os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r01"
GO, PR = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
from ncu_summary import WANT  # noqa: E402


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def summarize(rep, dst, links, units, unit_bytes, rows_per_unit):
    hdr, units_row, rows = raw(rep)
    r = rows[0]
    get = lambda k: r[hdr.index(k)] if k in hdr else None
    lines = [f"# ncu --set full --clock-control none, one launch; source: gpurun_out/{os.path.basename(rep)} (not tracked)",
             f"kernel: {get('Kernel Name')}", f"grid {get('Grid Size')} block {get('Block Size')}", ""]
    for w in WANT:
        if w in hdr:
            lines.append(f"{w:92s} {r[hdr.index(w)]:>18s} {units_row[hdr.index(w)]}")
    inst = float(get("smsp__inst_executed.sum").replace(",", ""))
    rd, wr = get("dram__bytes_read.sum"), get("dram__bytes_write.sum")
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}
    rdb = float(rd) * scale[units_row[hdr.index("dram__bytes_read.sum")]]
    wrb = float(wr) * scale[units_row[hdr.index("dram__bytes_write.sum")]]
    tick_rows = links * units * rows_per_unit
    lines += ["", f"derived: warp-instructions per 64-channel tick row = {inst / tick_rows:.2f}",
              f"derived: DRAM traffic per launch = {rdb + wrb:.0f} B (read {rdb:.0f} + write {wrb:.0f}); frame bytes = {links * units * unit_bytes}"]
    open(dst, "w").write("\n".join(lines) + "\n")
    return rdb + wrb


def instruction_buckets(rep, dst, link_units):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    isrc, iex, ist = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    data = [(r[isrc].strip(), int(r[iex]), int(r[ist])) for r in rows[2:] if len(r) > iex]
    tot, tst = sum(d[1] for d in data), sum(d[2] for d in data)
    ops = {}
    for s, e, _ in data:
        op = s.split()[1] if s.startswith("@") else s.split()[0]
        op = op.split(".")[0]
        ops[op] = ops.get(op, 0) + e
    lines = [f"# per-instruction counts from the same capture ({os.path.basename(rep)}); unit = executions per link-unit (one frame / superchunk of one link)",
             f"total warp-instructions per link-unit: {tot / link_units:.1f}   static SASS instructions: {len(data)}", "",
             "opcode            executed/link-unit   share"]
    for op, e in sorted(ops.items(), key=lambda kv: -kv[1])[:24]:
        lines.append(f"{op:16s} {e / link_units:14.1f} {100 * e / tot:8.1f}%")
    lines += ["", "by execution frequency (how often per link-unit each SASS line runs):"]
    b = {}
    for s, e, st in data:
        f = e / link_units
        k = ">=3.5 (tick loop, every 4-tick group)" if f >= 3.5 else "1.5-3.5 (per chunk)" if f >= 1.5 else "0.5-1.5 (busy tier)" if f >= 0.5 else "<0.5 (hit emission, link prologue/epilogue)"
        x = b.setdefault(k, [0, 0, 0])
        x[0] += e; x[1] += st; x[2] += 1
    for k, x in b.items():
        lines.append(f"  {k:48s} {x[0] / link_units:8.1f} instr/link-unit ({100 * x[0] / tot:4.1f}%)  stall samples {100 * x[1] / max(tst, 1):4.1f}%  static {x[2]}")
    open(dst, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    rep = os.path.join(GO, f"{tag}_wibeth_simple_full.ncu-rep")
    if os.path.exists(rep):
        traffic = summarize(rep, os.path.join(PR, f"{tag}_wibeth_simple_ncu_full.txt"), 5920, 64, 7200, 64)
        instruction_buckets(rep, os.path.join(PR, f"{tag}_wibeth_simple_instruction_mix.txt"), 5920 * 64)
        json.dump({"links": 5920, "frames": 64, "dram_bytes_per_launch": traffic,
                   "source": f"profiles/{tag}_wibeth_simple_ncu_full.txt (dram__bytes_read.sum + dram__bytes_write.sum, one launch)"},
                  open(os.path.join(PR, "dram_traffic.json"), "w"), indent=1)
    for name in ("fir", "absrs", "stdrs", "stress"):
        rep = os.path.join(GO, f"{tag}_wibeth_{name}_full.ncu-rep")
        if os.path.exists(rep):
            summarize(rep, os.path.join(PR, f"{tag}_wibeth_{name}_ncu_full.txt"), 5920, 64, 7200, 64)
            instruction_buckets(rep, os.path.join(PR, f"{tag}_wibeth_{name}_instruction_mix.txt"), 5920 * 64)
    for name in ("simple", "fir", "absrs"):
        rep = os.path.join(GO, f"{tag}_wib2_{name}_full.ncu-rep")
        if os.path.exists(rep):
            summarize(rep, os.path.join(PR, f"{tag}_wib2_{name}_ncu_full.txt"), 1480 * 4, 340, 5664 / 4, 12)
    src = os.path.join(GO, f"{tag}_launches.csv")
    if os.path.exists(src):
        keep = [l for l in open(src) if l.startswith('"')]
        open(os.path.join(PR, f"{tag}_launches.csv"), "w").writelines(keep)
