"""Per-source-line attribution of an ncu capture (development aid): joins the SASS rows of `ncu --page source --csv` with the
line table of `nvdisasm -gi` for the same kernel and sums executed warp-instructions and stall samples per CUDA source line.
usage: python tools/ncu_lines.py gpurun_out/x.ncu-rep <mangled kernel name> [units]   (units = warp-slots executed, to normalise)
needs the libdqlb200.so the capture ran with (compiled with -lineinfo)."""
import collections
import csv
import io
import pathlib
import re
import subprocess
import sys
import tempfile

ROOT = pathlib.Path(__file__).resolve().parent.parent
LIB = ROOT / "dql_multirotor_landing_b200" / "libdqlb200.so"


def line_table(kernel: str):
    """offset -> (file, line, sass) in execution order, inline call chain collapsed to the outermost train_kernel.cuh line too"""
    tmp = pathlib.Path(tempfile.mkdtemp())
    subprocess.run(["cuobjdump", "-xelf", "all", str(LIB)], cwd=tmp, check=True, capture_output=True)
    cubin = next(tmp.glob("*.cubin"))
    txt = subprocess.run(["nvdisasm", "-gi", "-c", str(cubin)], capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(txt) if l.startswith(".text." + kernel + ":"))
    rows, group, in_group = [], [("?", 0)], False
    for l in txt[start + 1:]:
        if l.startswith("\t.section") or l.startswith("//-----"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:      # consecutive markers: the inline chain of the next instruction, innermost location first
            loc = (pathlib.Path(m.group(1)).name, int(m.group(2)))
            group = group + [loc] if in_group else [loc]
            in_group = True
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            in_group = False
            # attribute to the innermost frame inside train_kernel.cuh (helpers of other headers are charged to their call site;
            # the body of a lambda to its own line, not to the line that calls the lambda)
            key = next((g for g in group if g[0] == "train_kernel.cuh"), group[-1])
            rows.append((int(m.group(1), 16), group[0], key, m.group(2).strip()))
    return rows


def main():
    rep, kernel = sys.argv[1], sys.argv[2]
    units = float(sys.argv[3]) if len(sys.argv) > 3 else 1.0
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    lines = raw.splitlines()
    # the first launch of the report only
    hdr = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
    end = next((i for i in range(hdr + 1, len(lines)) if lines[i].startswith('"Kernel Name"')), len(lines))
    rd = list(csv.DictReader(io.StringIO("\n".join(lines[hdr:end]))))
    tab = line_table(kernel)
    assert len(tab) == len(rd), (len(tab), len(rd))
    base = int(rd[0]["Address"], 16)
    stall_cols = [c for c in rd[0] if c.startswith("stall_") and "Not Issued" not in c]
    per_line = collections.defaultdict(lambda: collections.Counter())
    sass_out = []
    for r, (off, cur, outer, sass) in zip(rd, tab):
        assert int(r["Address"], 16) - base == off, (r["Address"], off)
        n = float(r["Instructions Executed"])
        key = outer
        c = per_line[key]
        c["inst"] += n
        c["samples"] += float(r["# Samples"])
        for s in stall_cols:
            c[s] += float(r[s])
        sass_out.append((off, cur, outer, sass, n, float(r["# Samples"]), {s: float(r[s]) for s in stall_cols if float(r[s]) > 0}))
    tot_i = sum(c["inst"] for c in per_line.values())
    tot_s = sum(c["samples"] for c in per_line.values())
    print(f"total warp-instructions {tot_i:.0f} ({tot_i / units:.1f} per unit), samples {tot_s:.0f}")
    print(f"{'file:line':34s} {'inst/unit':>9s} {'inst%':>6s} {'smp%':>6s}  top stalls")
    for key, c in sorted(per_line.items(), key=lambda kv: (kv[0][0], kv[0][1])):
        if c["inst"] / tot_i < 0.002 and c["samples"] / tot_s < 0.002:
            continue
        top = sorted(((s, c[s]) for s in stall_cols if c[s] > 0), key=lambda x: -x[1])[:3]
        print(f"{key[0] + ':' + str(key[1]):34s} {c['inst'] / units:9.1f} {100 * c['inst'] / tot_i:6.1f} {100 * c['samples'] / tot_s:6.1f}  " +
              ", ".join(f"{s[6:]} {100 * v / tot_s:.1f}" for s, v in top))
    if "--sass" in sys.argv:
        for off, cur, outer, sass, n, smp, st in sass_out:
            print(f"{off:05x} {outer[0]}:{outer[1]:<4d} {cur[0]}:{cur[1]:<4d} {n / units:7.2f} {100 * smp / tot_s:5.2f}  {sass[:70]:70s} " +
                  " ".join(f"{k[6:]}={v:.0f}" for k, v in sorted(st.items(), key=lambda x: -x[1])[:3]))


if __name__ == "__main__":
    main()
