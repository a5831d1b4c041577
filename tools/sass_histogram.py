"""SASS mnemonic histograms of the hot kernels of libserac_b200.so (cuobjdump -sass), for profiles/.

    cuobjdump -sass serac_b200/libserac_b200.so > /tmp/all.sass && python tools/sass_histogram.py /tmp/all.sass > profiles/<round>_sass_histograms.md

Selection: for each hot kernel the instance the bench step / config-A assembly launches (largest match by substring of the
demangled name), so the listing shows what the numbers in bench.py were measured on."""
import collections
import re
import subprocess
import sys

HOT = [
    ("evaluation_kernel (config A: SolidMechanics NeoHookean hex p=2, derivative store)", ["evaluation_kernel", "(serac::Geometry)5, 3, 0, serac::H1<2, 3>, serac::H1<1, 3>, serac::H1<2, 3>, serac::H1<2, 3> >, 1, serac::ShapeAwareIntegrandWrapperWithState<serac::MaterialStressFunctor<serac::solid_mechanics::NeoHookean, (serac::FlagMode)2>", "serac::Empty, 1, 16, true>"]),
    ("jvp_stream_kernel<3,3,3,3,3,3,24,2,5> (config A Jacobian action, symmetric tangent storage)", ["jvp_stream_kernel<3, 3, 3, 3, 3, 3, 24"]),
    ("scatter_kernel<false, true> (node-major E-vector)", ["scatter_kernel<false, true>"]),
    ("assemble_rows_kernel (config A fused assembly, symmetric tangent storage)", ["assemble_rows_kernel<3, 3, 3, 3, 3, 24"]),
    ("cg_spmv_dot_kernel", ["cg_spmv_dot_kernel"]),
    ("csr_spmv_kernel", ["csr_spmv_kernel"]),
]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def main(path):
    funcs, cur = {}, None
    arch = {}
    cur_arch = "?"
    with open(path) as f:
        for line in f:
            m = re.match(r"\s*arch = (\S+)", line)
            if m:
                cur_arch = m.group(1)
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                cur = m.group(1)
                funcs[cur] = collections.Counter()
                arch[cur] = cur_arch
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
            if m and cur:
                funcs[cur][m.group(1)] += 1
    names = demangle(list(funcs))
    print("# SASS mnemonic histograms of the hot kernels (`cuobjdump -sass serac_b200/libserac_b200.so`, tools/sass_histogram.py)\n")
    print(f"{len(funcs)} kernels in the library; cubin architectures: {sorted(set(arch.values()))}\n")
    for title, keys in HOT:
        cands = [m for m, d in names.items() if all(k in d for k in keys)]
        if not cands:
            cands = [m for m, d in names.items() if keys[0].split("<")[0] in d]
        if not cands:
            print(f"## {title}\n\nnot found\n")
            continue
        best = max(cands, key=lambda m: sum(funcs[m].values()))
        h = funcs[best]
        total = sum(h.values())
        groups = collections.Counter()
        for op, n in h.items():
            groups[op.split(".")[0]] += n
        print(f"## {title}\n")
        print(f"`{names[best][:300]}`  ({arch[best]}, {total} instructions)\n")
        print("| opcode | count | share |\n|---|---|---|")
        for op, n in groups.most_common(18):
            print(f"| {op} | {n} | {100.0 * n / total:.1f}% |")
        notable = {k: v for k, v in h.items() if k.startswith(("UBLKCP", "SYNCS", "LDGSTS", "DFMA", "DMUL", "DADD", "STG", "LDG", "LDS", "STS", "BAR", "UTMA", "FSEL", "SHFL"))}
        print("\nfull mnemonics of interest: " + ", ".join(f"{k} {v}" for k, v in sorted(notable.items(), key=lambda kv: -kv[1])[:24]) + "\n")


if __name__ == "__main__":
    main(sys.argv[1])
