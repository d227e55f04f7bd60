"""profiles/r02_sass_issue_loop.txt: SASS excerpts that prove the tcgen05 / TMEM / TMA path of the product
library (cuobjdump runs on the CPU build box; no GPU needed).   python scripts/sass_evidence.py"""
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "floodplanet_code_b200" / "lib" / "libfloodplanet_b200.so"
lines = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout.split("\n")
lines = [l for l in lines if l.strip()]


def func_range(name):
    starts = [i for i, l in enumerate(lines) if "Function :" in l]
    for j, i in enumerate(starts):
        if name in lines[i]:
            return i, (starts[j + 1] if j + 1 < len(starts) else len(lines))
    raise KeyError(name)


def strip(l):
    return re.sub(r"\s*/\* 0x[0-9a-f]+ \*/\s*$", "", l).rstrip()


MN = ["UTCHMMA", "LDTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "HMMA.16816", "HGMMA", "SYNCS", "UTCHMMA.WS", "FFMA2",
      "LDGSTS"]
out = ["Round 2 -- SASS evidence, product library floodplanet_code_b200/lib/libfloodplanet_b200.so (sm_100a)",
       "Command: cuobjdump -sass floodplanet_code_b200/lib/libfloodplanet_b200.so   (nvcc 12.9.86, -gencode "
       "arch=compute_100a,code=sm_100a -lineinfo)",
       "Regenerate: python scripts/sass_evidence.py", "",
       "Whole library mnemonic counts: " + ", ".join(f"{m} x{sum(m in l for l in lines)}" for m in MN),
       "(UTCHMMA = tcgen05.mma kind::f16; LDTM = tcgen05.ld; UTMALDG / UTMASTG = cp.async.bulk.tensor load / store (TMA);",
       " UTCBAR = tcgen05.commit -> mbarrier; UTCATOMSWS = tcgen05.alloc / dealloc; no legacy HMMA.16816 (mma.sync) and no",
       " HGMMA (wgmma) anywhere in the library; UTCHMMA.WS = tcgen05.mma.ws with collector qualifiers, FFMA2 = fma.rn.f32x2,",
       " LDGSTS = cp.async -- the last two in the head backward)", ""]
KERNELS = [
    ("fprop / dgrad: conv3x3_halo_kernel<BN=128, KCH=64, TAPS=9, EW=4>  (the dominant kernel of the step)",
     "conv3x3_halo_kernelILi128ELi64ELi9ELi4E"),
    ("fprop / dgrad: conv3x3_halo_kernel<BN=64, KCH=64, TAPS=9, EW=8, WS=true>  (Cout = 64 full-resolution layers: "
     "weight-stationary UTCHMMA.WS, B_KEEP = collector fill by tile 0, B_REUSE = last use by tile 1, BUFFER<k> = K slice)",
     "conv3x3_halo_kernelILi64ELi64ELi9ELi8ELb1E"),
    ("wgrad: conv3x3_wgrad_kernel<MODE_X_SHIFT=0, NBW=64, NB=2>  (Cout >= 128 layers: three taps stacked on N = 192)",
     "conv3x3_wgrad_kernelILi0ELi64ELi2E"),
    ("wgrad: conv3x3_wgrad_kernel<MODE_RS_SPLIT=4, NBW=64, NB=1>  (Cout = 64 layers: rows stacked on M, taps on N)",
     "conv3x3_wgrad_kernelILi4ELi64ELi1E"),
    ("wgrad: conv3x3_wgrad_kernel<MODE_X_STACK=3, NBW=16, NB=1>  (first layer, Cin 4 -> 16)",
     "conv3x3_wgrad_kernelILi3ELi16ELi1E"),
]
for title, name in KERNELS:
    a, b = func_range(name)
    body = lines[a:b]
    per = {m: sum(m in l for l in body) for m in MN[:5]}
    out += ["=" * 118, title, lines[a].strip(), f"per-kernel counts: {per}", ""]
    idx = [i for i, l in enumerate(body) if "UTCHMMA" in l]
    lo, hi = max(0, idx[0] - 4), idx[min(11, len(idx) - 1)] + 3
    out.append(f"-- MMA issue loop (first {min(12, len(idx))} of {len(idx)} UTCHMMA; only uniform-datapath descriptor "
               "arithmetic in between):")
    out += [strip(l) for l in body[lo:hi]]
    for m, ctx in (("UTMALDG", 1), ("UTMASTG", 1), ("LDTM", 1)):
        ids = [i for i, l in enumerate(body) if m in l][:2]
        if not ids:
            continue
        out.append(f"-- {m} sites (first {len(ids)} of {sum(m in l for l in body)}):")
        for i in ids:
            out += [strip(l) for l in body[max(0, i - ctx):i + ctx + 1]]
            out.append("        ...")
    out.append("")
(ROOT / "profiles" / "r02_sass_issue_loop.txt").write_text("\n".join(out) + "\n")
print("wrote profiles/r02_sass_issue_loop.txt", len(out), "lines")
