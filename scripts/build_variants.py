"""Build variants of libpgc.so that differ only in compile-time switches of one source file (kernel experiments; default
eval_cec2014.cu, another one with --src=eval_lj.cu as the first argument).

    python scripts/build_variants.py name1:-DPGC_WARPS=20,-DPGC_ZT_SWIZZLE=1 name2:-DPGC_GEMM_UNROLL=5 ...

Output: pagmo2_b200/_variants/libpgc_<name>.so (git-ignored; travels to the GPU box).  Select one at run time with
PGC_LIBRARY_PATH=pagmo2_b200/_variants/libpgc_<name>.so.
"""
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from pagmo2_b200 import build as B  # noqa: E402

B.build()
out = B.HERE / "_variants"
out.mkdir(exist_ok=True)
args = sys.argv[1:]
src_name = "eval_cec2014.cu"
if args and args[0].startswith("--src="):
    src_name = args.pop(0)[6:]
src = B.CSRC / src_name
stem = src_name.rsplit(".", 1)[0]
for spec in args:
    name, _, flags = spec.partition(":")
    flags = [f for f in flags.split(",") if f]
    obj = out / f"{stem}_{name}.o"
    cmd = [B.NVCC, *B.ARCH, *B.COMMON, *flags, "-x", "cu", "-c", str(src), "-o", str(obj)]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode:
        sys.exit(r.stderr)
    info = [l for l in r.stderr.splitlines() if "Used" in l or "spill" in l]
    lines = r.stderr.splitlines()
    for i, l in enumerate(lines):
        if ("stage_kernelILi100ELb1" in l or "lj_circ_kernelILi5" in l) and "Compiling" in l:
            print(name, " | ".join(x.strip() for x in lines[i + 2:i + 4]))
    objs = [str(o) for o in sorted(B.OBJDIR.glob("*.o")) if o.name != src_name + ".o"] + [str(obj)]
    so = out / f"libpgc_{name}.so"
    r = subprocess.run([B.NVCC, *B.ARCH, "-shared", "-o", str(so), *objs, "-Xcompiler", "-fPIC", "-cudart", "static"],
                       capture_output=True, text=True)
    if r.returncode:
        sys.exit(r.stderr)
    obj.unlink()
    print("built", so)
