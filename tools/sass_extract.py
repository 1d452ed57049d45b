"""Extract cuobjdump -sass listings of the hot kernels from the shipped libtrikb200.so into profiles/ (gzip) and an opcode
histogram per kernel (text).  Runs without a GPU.  usage: python tools/sass_extract.py <tag>   e.g. r02a"""
import collections, gzip, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "trik-media-sensors-dsp_b200", "libtrikb200.so")
# the instantiations the bench / sweeps actually launch
WANT = ["vsum_kernelILb0ELi4ELi512E", "vsum16_kernelILi4ELi256ELi4E", "wo_lut_kernelILi4ELb0E", "om_table_kernelE",
        "om_table_list_kernelE", "om_major_kernelILi4ELi3E", "oo_bitmap_lut_kernelILb0E", "oo_cluster_kernelE",
        "preview_identity_kernelILi1ELb1E", "preview_identity_kernelILi3ELb1E", "preview_identity_kernelILi1ELb0E"]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r02"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    parts = re.split(r"(?m)^\s*Function : ", sass)
    out_sass, summary = [], []
    for p in parts[1:]:
        name = p.split("\n", 1)[0].strip()
        if not any(w in name for w in WANT):
            continue
        out_sass.append("Function : " + p)
        ops = collections.Counter()
        for m in re.finditer(r"(?m)^\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", p):
            ops[m.group(1).split(".")[0] + ("." + m.group(1).split(".")[1] if m.group(1).startswith(("IDP", "VIADD", "VIMNMX", "LDG", "LDGSTS", "ATOMS", "REDUX", "MATCH")) and "." in m.group(1) else "")] += 1
        total = sum(ops.values())
        summary.append("%s\n  %d instructions; %s\n" % (name, total, ", ".join("%s %d" % kv for kv in ops.most_common(22))))
    with gzip.open(os.path.join(ROOT, "profiles", tag + "_sass_hot_kernels.txt.gz"), "wt") as f:
        f.write("\n".join(out_sass))
    with open(os.path.join(ROOT, "profiles", tag + "_sass_hot_kernels_opcodes.txt"), "w") as f:
        f.write("cuobjdump -sass of libtrikb200.so (sm_100a), static opcode counts of the hot kernels' instantiations the bench launches\n"
                "(full listings: %s_sass_hot_kernels.txt.gz).  No UTCMMA / LDTM / UTMALDG by design: nothing here is a contraction,\n"
                "rows are staged by the per-thread LDGSTS (cp.async) ring (DESIGN 3.1).\n\n" % tag)
        f.write("\n".join(summary))
    print("".join(summary))


if __name__ == "__main__":
    main()
