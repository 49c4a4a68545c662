"""Debug driver: bf16 gradient GEMMs in isolation (one subprocess per case)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import conftest  # noqa


def run(case):
    import torch
    from test_gpu_train_kernels import conv_grads_case
    if case == "dgrad1x1":
        print(case, conv_grads_case(2, 16, 16, 64, 128, 1, 1, gdt=torch.bfloat16))
    elif case == "dgrad3x3":
        print(case, conv_grads_case(2, 16, 16, 64, 128, 3, 1, gdt=torch.bfloat16))
    elif case == "f16":
        print(case, conv_grads_case(2, 16, 16, 64, 128, 3, 1))
    torch.cuda.synchronize()


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1])
    else:
        for c in ("f16", "dgrad1x1", "dgrad3x3"):
            r = subprocess.run([sys.executable, __file__, c], capture_output=True, text=True)
            print("==", c, "rc", r.returncode, (r.stdout.strip().splitlines() or [""])[-1], "|",
                  " ".join((r.stderr.strip().splitlines() or [""])[-2:])[:300], flush=True)
