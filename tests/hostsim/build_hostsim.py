"""TEST INFRASTRUCTURE ONLY: builds tests/hostsim/_build/libhostsim.so (device headers compiled for the CPU)."""
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "svb_models_asl_b200", "csrc")
OUT = os.path.join(HERE, "_build")
LIB = os.path.join(OUT, "libhostsim.so")

sys.path.insert(0, ROOT)


def build(defines=()):
    """defines: extra -D switches of the device headers (e.g. ("SVB_SUM_FORM=0",)): a variant library beside the
    default one, for tests that compare two formulations of the same arithmetic."""
    from svb_models_asl_b200.build import aslrest_flag_sets, disp_flag_sets, variants
    os.makedirs(OUT, exist_ok=True)
    tag = "".join("_" + d.replace("=", "").lower() for d in defines)
    lib = os.path.join(OUT, "libhostsim%s.so" % tag)
    flags = aslrest_flag_sets()
    fast = sorted({(f, nbt) for (_m, kind, f, nbt, _l, _e) in variants() if kind == 0 and nbt})
    hdr = ["// GENERATED", "#define HOSTSIM_ASLREST_FLAGS " + " ".join("X(0x%xu)" % f for f in flags),
           "#define HOSTSIM_DISP_FLAGS " + " ".join("X(0x%xu)" % f for f in disp_flag_sets()),
           "#define HOSTSIM_FAST " + " ".join("Y(0x%xu, %d)" % v for v in fast), ""]
    text = "\n".join(hdr)
    path = os.path.join(OUT, "model_list.h")
    if not os.path.exists(path) or open(path).read() != text:
        open(path, "w").write(text)
    srcs = [os.path.join(HERE, "hostsim.cpp"), path] + [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))
                                                       if f.endswith(".h")]
    srcs.append(os.path.join(ROOT, "include", "svbasl.h"))
    h = hashlib.sha256()
    for s in srcs:
        h.update(open(s, "rb").read())
    h.update(" ".join(defines).encode())
    stamp = os.path.join(OUT, "digest%s.txt" % tag)
    if os.path.exists(lib) and os.path.exists(stamp) and open(stamp).read() == h.hexdigest():
        return lib
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-I", CSRC, "-I", OUT] + \
          ["-D" + d for d in defines] + [os.path.join(HERE, "hostsim.cpp"), "-o", lib]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("hostsim build failed:\n" + res.stderr[-6000:])
    open(stamp, "w").write(h.hexdigest())
    return lib


if __name__ == "__main__":
    print(build())
