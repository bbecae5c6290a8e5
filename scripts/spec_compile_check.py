"""Developer tool: is a model eligible for an ahead-of-time specialised kernel, and does the kernel compile for it?
    python scripts/spec_compile_check.py <robot.urdf> <plane.urdf|none> <floating 0|1> <n_act> <start_link>
Compiles the URDF with our model compiler, runs csrc/gen_spec.cpp on the flat model and instantiates
tds_step_spec_kernel<Spec, ...> for sm_100a (all three arithmetics, all three instances).  No GPU needed."""
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from tds_b200.model import compile_urdf  # noqa: E402

CSRC = os.path.join(ROOT, "tiny-differentiable-simulator_b200", "csrc")
INC = os.path.join(ROOT, "include")


def main():
    urdf, plane, floating, n_act, start = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4], sys.argv[5]
    m = compile_urdf(urdf, None if plane == "none" else plane, bool(floating))
    with tempfile.TemporaryDirectory() as d:
        inc = os.path.join(d, "model.inc")
        with open(inc, "w") as f:
            f.write("// flat model\n")
            vals = [repr(float(v)) for v in m]
            for i in range(0, len(vals), 6):
                f.write(", ".join(vals[i:i + 6]) + ",\n")
        exe = os.path.join(d, "gen_spec")
        subprocess.check_call([os.environ.get("TDS_CXX", "/usr/bin/g++"), "-std=c++17", "-O1", "-I", CSRC, "-I", INC,
                               os.path.join(CSRC, "gen_spec.cpp"), "-o", exe])
        subprocess.check_call([exe, "SpecTest", inc, n_act, start, os.path.join(d, "spec_test.h")])
        cu = os.path.join(d, "check.cu")
        with open(cu, "w") as f:
            f.write('#include "spec_test.h"\n#include "tds_steps.cu"\nTDS_SPEC_TABLES(SpecTest, test)\n'
                    'extern "C" int spec_test_launch(const SimParams* P, const EnvParams* E, const StepIO* io, int mode, int use_pd,\n'
                    '                                int precision, cudaStream_t st) {\n'
                    '  return tdss::SpecHost<SpecTest>::launch(P, E, io, mode, use_pd, precision, st);\n}\n'
                    'extern "C" size_t spec_test_smem(int p) { return tdss::SpecHost<SpecTest>::smem_bytes(p); }\n')
        r = subprocess.run([os.environ.get("TDS_NVCC", "/usr/local/cuda/bin/nvcc"), "-gencode", "arch=compute_100a,code=sm_100a", "-O3",
                            "-std=c++17", "-diag-suppress", "549,177", "-I", d, "-I", CSRC, "-I", INC, "-Xptxas", "-v", "-c", cu,
                            "-o", os.path.join(d, "check.o")], capture_output=True, text=True)
        if r.returncode:
            print(r.stderr[-4000:])
            print("NOT ELIGIBLE / does not compile (the subtrees must be one structural class: see Cls<SP>::uniform)")
            return 1
        regs = [ln for ln in r.stderr.splitlines() if "registers" in ln]
        print("compiles: %d kernel instances; e.g. %s" % (len(regs), regs[-1].strip() if regs else ""))
    return 0


if __name__ == "__main__":
    sys.exit(main())
