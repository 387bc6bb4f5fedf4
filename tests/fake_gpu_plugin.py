"""TEST INFRASTRUCTURE: `python -m pytest -p fake_gpu_plugin tests/<file> -m gpu` dry-runs device tests in a container without a GPU by
swapping MapleEngine for tests/fake_device.FakeEngine (the oracle and the host-compiled CUDA source behind the ctypes table).  It
checks the test harness and the python orchestration, not the kernels; used before sending new device tests to the hardware.  Never
active unless asked for with -p."""


def pytest_configure(config):
    import torch

    import maple_b200.engine as engine
    from fake_device import FakeEngine

    def make(model, device=0):
        return FakeEngine(model)

    engine.MapleEngine = make
    torch.cuda.synchronize = lambda *a, **k: None
