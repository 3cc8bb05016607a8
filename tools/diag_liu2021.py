import sys, os
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/comprehensive-transformer-tts_b200"); sys.path.insert(0, "/root/repo/tests")
sys.path.insert(0, "/root/repo/tests/golden")
import test_gpu_configs as T
for M in (64, 256, 1024):
    try:
        T.test_config4_fs2_liu2021_length_sweep(M)
    except AssertionError as e:
        print("ASSERT", str(e)[:200])
