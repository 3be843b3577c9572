import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")
    config.addinivalue_line("markers", "reference: needs the read-only reference tree at /root/reference")


# GPU cases added after the last B200 call of round 1 (the round's GPU budget was spent): they are parity-green on the
# 1-lane emulation of the same device source and against the oracle / reference, but have not run on a B200 yet.
# They run LAST, so that with `-x` a surprise there cannot mask the cases that are known to pass on the hardware.
NOT_YET_RUN_ON_B200 = ("full_obs", "compact_transfer", "[change_list]", "lognormal_reset", "split_reset", "c3_full_size", "c5_full_size", "covid_cuda_full_size", "reference_api_cuda", "us_federal_annealed")   # (everything else passed on a B200: profiles/r01f_pytest_gpu.log)


def pytest_collection_modifyitems(config, items):
    import torch

    has_gpu = torch.cuda.is_available()
    for item in items:
        if "gpu" in item.keywords and not has_gpu:
            item.add_marker(pytest.mark.skip(reason="no CUDA device"))
    late = [i for i in items if "gpu" in i.keywords and any(n in i.nodeid for n in NOT_YET_RUN_ON_B200)]
    if late:
        ids = {id(i) for i in late}
        items[:] = [i for i in items if id(i) not in ids] + late
