"""Regenerate profiles/sass/ (mnemonic histograms of every object + full listings of the tensor-core kernels).
Runs on the CPU box: `python bench/sass_summary.py` after `python __graft_entry__.py`."""
import collections
import glob
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = glob.glob(os.path.join(ROOT, "heterofl*_b200"))[0]
OBJ = os.path.join(PKG, "ops", "_build")
OUT = os.path.join(ROOT, "profiles", "sass")
KEYS = ["UTCHMMA", "UTCQMMA", "UTCCP", "HMMA", "UTMALDG", "UTMALDG.2D.MULTICAST", "UBLKCP", "LDTM", "UTCBAR", "UTCBAR.MULTICAST",
        "UTCATOMSWS", "UCGABAR_ARV", "UCGABAR_WAIT", "MAPA", "ACQBULK", "PREEXIT", "SYNCS", "ATOMS", "ATOMG", "RED", "BAR", "LDG", "STG",
        "LDS", "STS", "SHFL", "MUFU", "HFMA2"]


def sass(obj, fun=None):
    cmd = ["cuobjdump", "-sass"] + (["-fun", fun] if fun else []) + [obj]
    return subprocess.run(cmd, capture_output=True, text=True).stdout


def main():
    os.makedirs(OUT, exist_ok=True)
    lines = ["# SASS evidence (cuobjdump -sass of the in-tree objects, sm_100a)", "",
             "Mnemonic histogram per translation unit (regenerate with `python bench/sass_summary.py`).  `UTCHMMA` = tcgen05.mma",
             "(kind::f16), `UTMALDG` = TMA cp.async.bulk.tensor (`.MULTICAST` = cluster multicast), `LDTM` = tcgen05.ld,",
             "`UTCBAR` = tcgen05.commit -> mbarrier (`.MULTICAST` = arrive on the same barrier in every CTA of the mask),",
             "`SYNCS` = mbarrier ops, `UCGABAR_ARV` = cluster barrier, `ATOMS`/`RED`/`ATOMG` = shared/global atomics; `LDG.E...` on",
             "peer-mapped pointers are the P2P loads of `aggregate_kernel`.  `UTCQMMA` / `UTCCP` / `UBLKCP` = block-scaled",
             "tcgen05.mma (MXFP8), tcgen05.cp of the scale factors, cp.async.bulk; `HMMA` = mma.sync (attention kernels);",
             "`ACQBULK` / `PREEXIT` = griddepcontrol.wait / launch_dependents (programmatic dependent launch, every kernel);",
             "`UCGABAR_ARV` / `UCGABAR_WAIT` + `MAPA` + `LD...` on `shared::cluster` = the cluster split-K reduction of the",
             "convolution and the cluster BatchNorm kernels.  Full listings: `gemm_mc_tcgen05.sass` (cluster + multicast GEMM,",
             "all variants), `gemm_tcgen05.sass` (one TN and one general instantiation), `conv_tcgen05.sass` (fprop<128> with",
             "residual + statistics, incl. the split-K path; dgrad<128>; wgrad), `lm_vocab_tcgen05.sass` (fused vocabulary",
             "head, backward kernel), `lm_attn_mma.sass` (tensor-core attention).", ""]
    for obj in sorted(glob.glob(os.path.join(OBJ, "*.cu.o"))):
        txt = sass(obj)
        kernels = re.findall(r"Function : (\S+)", txt)
        cnt = collections.Counter()
        total = 0
        for ln in txt.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", ln)
            if not m:
                continue
            total += 1
            op = m.group(1)
            base = op.split(".")[0]
            cnt[base] += 1
            if op.startswith("UTMALDG") and "MULTICAST" in op:
                cnt["UTMALDG.2D.MULTICAST"] += 1
            if op.startswith("UTCBAR") and "MULTICAST" in op:
                cnt["UTCBAR.MULTICAST"] += 1
        lines += ["## %s" % os.path.basename(obj), "", "%d kernels" % len(kernels), "", "| mnemonic | count |", "|---|---|"]
        lines += ["| %s | %d |" % (k, cnt[k]) for k in KEYS if cnt[k]]
        lines += ["| (total instructions) | %d |" % total, ""]
    with open(os.path.join(OUT, "README.md"), "w") as f:
        f.write("\n".join(lines))

    def pick(obj, pats):
        txt = sass(os.path.join(OBJ, obj))
        names = re.findall(r"Function : (\S+)", txt)
        out = []
        for p in pats:
            for n in names:
                if re.search(p, n):
                    out.append(sass(os.path.join(OBJ, obj), n))
                    break
        return "\n".join(out)
    open(os.path.join(OUT, "gemm_mc_tcgen05.sass"), "w").write(sass(os.path.join(OBJ, "gemm_mc_tcgen05.cu.o")))
    open(os.path.join(OUT, "gemm_tcgen05.sass"), "w").write(
        pick("gemm_tcgen05.cu.o", [r"gemm_tcgen05_kernelILi128ELb1ELi2ELb0", r"gemm_tcgen05_general_kernelILi128ELb1ELb1ELb1"]))
    open(os.path.join(OUT, "conv_tcgen05.sass"), "w").write(
        pick("conv_tcgen05.cu.o", [r"conv_fprop_tcgen05_kernelILi128ELb0ELb1ELb1", r"conv_fprop_tcgen05_kernelILi128ELb1ELb0ELb0",
                                   r"conv_wgrad_tcgen05_kernel"]))
    open(os.path.join(OUT, "lm_vocab_tcgen05.sass"), "w").write(pick("lm_vocab_tcgen05.cu.o", [r"lm_vocab_bwd_kernel"]))
    open(os.path.join(OUT, "lm_attn_mma.sass"), "w").write(sass(os.path.join(OBJ, "lm_attn_mma.cu.o")))


if __name__ == "__main__":
    main()
