#!/usr/bin/env python
"""SASS evidence for libcpg.so (north star: "kernel choices are evidenced by committed SASS").

    python tools/sass_report.py [profiles/r02]

Reads `cuobjdump -sass curdleproofs_pie_b200/lib/libcpg.so` and writes
  <prefix>_sass_histogram.txt   per kernel: instruction count, the multiply-pipe mix (IMAD.WIDE.U32[.X] = one 32x32->64
                                multiply-accumulate with carry; IMAD/IMAD.HI narrow), adds, memory, shuffles, calls, and
                                the register-ABI subroutines it calls (Fq product / square bodies) with their own mix
  <prefix>_sass_fq_mul_sqr.txt  the full SASS of the Fq product and Fq square subroutines of BucketAccumulate
It also prints a JSON summary (used by tests/test_abi_exports.py::test_sass_is_sm100a_integer_pipe)."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "curdleproofs_pie_b200", "lib", "libcpg.so")
INS = re.compile(r"^\s+/\*([0-9a-f]{4,})\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)\s*(.*?);")
TENSOR_OR_TMA = ("HMMA", "IMMA", "DMMA", "QMMA", "OMMA", "UTCHMMA", "UTCIMMA", "UTCQMMA", "UTCOMMA", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "LDSM")


def demangle_kernel(name):
    m = re.search(r"k_eachIN(?:S_|3cpg)\d*([A-Za-z0-9]+?)ELi(\d+)ELi(\d+)E", name)
    if m:
        return "%s<%s,%s>" % (m.group(1).lstrip("0123456789"), m.group(2), m.group(3))
    m = re.search(r"\d+(k_[a-z_0-9]+?)E", name)
    return m.group(1) if m else name


def parse(text):
    """-> {kernel: [(addr, opcode, operands)]}"""
    out, cur = collections.OrderedDict(), None
    for line in text.splitlines():
        if "Function :" in line:
            cur = demangle_kernel(line.split("Function :")[1].strip())
            out[cur] = []
            continue
        m = INS.match(line)
        if m and cur is not None:
            out[cur].append((int(m.group(1), 16), m.group(2), m.group(3)))
    return out


def mix(ins):
    c = collections.Counter()
    for _, op, _ in ins:
        if op.startswith("IMAD.WIDE.U32"):
            c["wide"] += 1
        elif op.startswith("IMAD.WIDE"):
            c["wide_signed"] += 1
        elif op.startswith("IMAD.MOV") or op == "IMAD.U32" and False:
            c["imad_mov"] += 1
        elif op.startswith("IMAD"):
            c["imad_narrow"] += 1
        elif op.startswith("IADD3") or op.startswith("IADD") or op.startswith("UIADD"):
            c["iadd"] += 1
        elif op.startswith(("LDG", "STG", "LD.", "ST.")) or op in ("LD", "ST"):
            c["global_mem"] += 1
        elif op.startswith(("LDL", "STL")):
            c["local_mem"] += 1
        elif op.startswith(("LDS", "STS")):
            c["shared_mem"] += 1
        elif op.startswith("SHFL"):
            c["shfl"] += 1
        elif op.startswith("CALL"):
            c["call"] += 1
        elif op.startswith(("LOP3", "SHF", "PRMT", "SEL", "ISETP", "PLOP3", "MOV", "UMOV", "ULOP", "USHF")):
            c["alu_other"] += 1
        if op.startswith(TENSOR_OR_TMA):
            c["tensor_or_tma"] += 1
        c["total"] += 1
    return c


def subroutines(ins):
    """bodies reached by CALL.REL.NOINC <addr>: from the target up to the first RET"""
    targets = collections.Counter()
    for _, op, args in ins:
        if op.startswith("CALL"):
            m = re.search(r"0x([0-9a-f]+)", args)
            if m:
                targets[int(m.group(1), 16)] += 1
    by_addr = {a: i for i, (a, _, _) in enumerate(ins)}
    subs = []
    for t, ncalls in sorted(targets.items()):
        if t not in by_addr:
            continue
        i = by_addr[t]
        body = []
        while i < len(ins):
            body.append(ins[i])
            if ins[i][1].startswith("RET"):
                break
            i += 1
        subs.append((t, ncalls, body))
    return subs


def classify(body_mix):
    w = body_mix["wide"]
    if 270 <= w <= 300:
        return "Fq product (12-limb even/odd CIOS: 288 product+reduction MACs + 12 for m_i)"
    if 200 <= w <= 230:
        return "Fq square (78 product MACs + 144 reduction MACs)"
    if w > 300:
        return "run of Fq squarings / larger field routine"
    return "helper"


def main():
    prefix = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02")
    text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", text)))
    kernels = parse(text)
    lines = ["# cuobjdump -sass curdleproofs_pie_b200/lib/libcpg.so   (arch: %s; %d kernels)" % (", ".join(arch), len(kernels)),
             "# wide = IMAD.WIDE.U32[.X] (one 32x32->64 MAC with carry-in/out: the roofline's unit, 32 lanes/clk/SM on B200)",
             "# narrow = other IMAD (32-bit multiply-add / address arithmetic; IMAD.MOV counted separately)",
             "%-28s %8s %8s %7s %7s %7s %7s %7s %7s %6s %6s %6s" % ("kernel", "instrs", "wide", "wide%", "narrow", "iadd", "gmem", "lmem", "smem", "shfl", "call", "mma/tma")]
    tot = collections.Counter()
    summary = {"arch": arch, "kernels": {}}
    sub_dump = []
    for k, ins in kernels.items():
        c = mix(ins)
        tot.update(c)
        lines.append("%-28s %8d %8d %6.1f%% %7d %7d %7d %7d %7d %6d %6d %6d" % (
            k, c["total"], c["wide"], 100.0 * c["wide"] / max(1, c["total"]), c["imad_narrow"], c["iadd"], c["global_mem"], c["local_mem"], c["shared_mem"], c["shfl"], c["call"], c["tensor_or_tma"]))
        summary["kernels"][k] = {"instrs": c["total"], "wide": c["wide"], "tensor_or_tma": c["tensor_or_tma"], "local_mem": c["local_mem"]}
        for t, ncalls, body in subroutines(ins):
            bm = mix(body)
            if bm["wide"] >= 100:
                lines.append("    sub @0x%04x  called from %3d sites  %5d instrs  wide %4d  narrow %3d  iadd %3d  mem %d   %s" % (
                    t, ncalls, bm["total"], bm["wide"], bm["imad_narrow"], bm["iadd"], bm["global_mem"] + bm["local_mem"], classify(bm)))
                if k == "BucketAccumulate<128,3>":
                    sub_dump.append((t, ncalls, body, bm))
    lines.append("%-28s %8d %8d %6.1f%% %7d %7d %7d %7d %7d %6d %6d %6d" % (
        "ALL", tot["total"], tot["wide"], 100.0 * tot["wide"] / max(1, tot["total"]), tot["imad_narrow"], tot["iadd"], tot["global_mem"], tot["local_mem"], tot["shared_mem"], tot["shfl"], tot["call"], tot["tensor_or_tma"]))
    summary["total"] = {"instrs": tot["total"], "wide": tot["wide"], "tensor_or_tma": tot["tensor_or_tma"]}
    os.makedirs(os.path.dirname(prefix), exist_ok=True)
    with open(prefix + "_sass_histogram.txt", "w") as f:
        f.write("\n".join(lines) + "\n")
    with open(prefix + "_sass_fq_mul_sqr.txt", "w") as f:
        f.write("# Fq product / Fq square subroutines as linked into BucketAccumulate<128,3> (register ABI: operands and result in\n"
                "# registers, no stack traffic; -DCPG_FIELD_CALLS).  Source: csrc/bigint.cuh mont_mul_n / mont_sqr_n.\n")
        for t, ncalls, body, bm in sub_dump:
            f.write("\n# ---- sub @0x%04x: %s; %d instrs, %d IMAD.WIDE.U32[.X], called from %d sites\n" % (t, classify(bm), bm["total"], bm["wide"], ncalls))
            for a, op, args in body:
                f.write("  /*%04x*/  %-22s %s ;\n" % (a, op, args))
    summary["bucket_accumulate_subs"] = [{"addr": t, "wide": bm["wide"], "instrs": bm["total"], "mem": bm["global_mem"] + bm["local_mem"]} for t, _, _, bm in sub_dump]
    print(json.dumps(summary))


if __name__ == "__main__":
    main()
