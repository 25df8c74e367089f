"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (count, total, mean, share)."""
import csv
import re
import sys
from collections import OrderedDict


def main(path, skip_setup=True):
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    for r in rd:
        if len(r) <= iv:
            continue
        name = re.sub(r"\(.*", "", r[ik]).replace("<unnamed>::", "")
        name = re.sub(r"void |cub::CUB_\d+_NS::", "", name)
        name = re.sub(r"<.*", "", name)
        rows.append((name, float(r[iv].replace(",", ""))))
    setup = {"split_linv_kernel", "copy_pad_kernel", "gemm_nn_f64_kernel", "tri_diag_inverse_kernel", "tri_embed_kernel", "tri_extract_kernel"}
    agg = OrderedDict()
    for name, ns in rows:
        if skip_setup and name in setup:
            continue
        c, t = agg.get(name, (0, 0.0))
        agg[name] = (c + 1, t + ns)
    tot = sum(t for _, t in agg.values())
    print(f"# {path}: {len(rows)} launches, {tot / 1e6:.3f} ms of kernel time (once-per-epoch setup kernels excluded: {sorted(setup)})")
    print(f"{'kernel':48s} {'launches':>8s} {'total ms':>10s} {'mean us':>10s} {'share':>7s}")
    for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{name[:48]:48s} {c:8d} {t / 1e6:10.3f} {t / c / 1e3:10.2f} {100 * t / tot:6.2f}%")
    su = [(n, ns) for n, ns in rows if n in setup]
    if su:
        print("# setup kernels (dmo_gp_create, once per epoch):")
        for n in sorted(setup):
            ts = [ns for m, ns in su if m == n]
            if ts:
                print(f"{n:48s} {len(ts):8d} {sum(ts) / 1e6:10.3f} {sum(ts) / len(ts) / 1e3:10.2f}")


if __name__ == "__main__":
    main(sys.argv[1])
