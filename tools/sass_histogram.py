"""SASS opcode histogram of one kernel of libdqlb200.so (static counts from `nvdisasm`; with an ncu capture also the executed
warp-instructions per opcode).  Evidence for the data-movement path: UBLKCP = cp.async.bulk (the copy engine), SYNCS = mbarrier
operations, LDGSTS = per-thread cp.async, LDG / STG = plain global loads / stores.
usage: python tools/sass_histogram.py <mangled kernel> out.json [capture.ncu-rep units]"""
import collections
import json
import pathlib
import subprocess
import sys

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent))
import ncu_lines


def opcode(sass: str) -> str:
    tok = sass.split()
    op = tok[1] if tok[0].startswith("@") else tok[0]
    return op.split(".")[0]


def main():
    kernel, out = sys.argv[1], sys.argv[2]
    tab = ncu_lines.line_table(kernel)
    static = collections.Counter(opcode(s) for _, _, _, s in tab)
    full = collections.Counter(".".join((s.split()[1] if s.split()[0].startswith("@") else s.split()[0]).split(".")[:3]) for _, _, _, s in tab)
    res = {"kernel": kernel, "instructions_static": len(tab), "opcodes_static": dict(static.most_common()),
           "data_movement_static": {k: v for k, v in sorted(full.items()) if k.split(".")[0] in ("UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDGSTS", "LDG", "STG", "FENCE", "LDGDEPBAR")}}
    if len(sys.argv) > 4:
        import csv, io
        rep, units = sys.argv[3], float(sys.argv[4])
        raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout.splitlines()
        hdr = next(i for i, l in enumerate(raw) if l.startswith('"Address"'))
        end = next((i for i in range(hdr + 1, len(raw)) if raw[i].startswith('"Kernel Name"')), len(raw))
        rows = list(csv.DictReader(io.StringIO("\n".join(raw[hdr:end]))))
        assert len(rows) == len(tab)
        ex = collections.Counter()
        for r, (_, _, _, s) in zip(rows, tab):
            ex[opcode(s)] += float(r["Instructions Executed"])
        res["executed_per_unit"] = {k: round(v / units, 2) for k, v in ex.most_common() if v / units >= 0.05}
        res["executed_total_per_unit"] = round(sum(ex.values()) / units, 1)
        res["unit"] = "one warp-slot = 32 env-steps (capture: %s, %d units)" % (pathlib.Path(rep).name, int(units))
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps({k: res[k] for k in ("instructions_static", "data_movement_static")}, indent=1))


if __name__ == "__main__":
    main()
