"""A stand-in for the handful of torch calls bench.py makes -- TEST INFRASTRUCTURE (tests/emu/run_bench_emu.py only).

It lets the benchmark driver's control flow, JSON contract and arithmetic be smoke-tested on a machine without a GPU, on top
of the emulated kernels.  Timings it produces are host wall-clock times of the emulation and mean nothing."""
import time

import numpy as np

uint8 = np.uint8
int16 = np.int16
float64 = np.float64


class _Tensor:
    def __init__(self, arr):
        self.arr = arr

    def fill_(self, v):
        self.arr.fill(v)
        return self

    def copy_(self, other):
        self.arr[...] = other.arr.astype(self.arr.dtype)
        return self

    def pin_memory(self):
        return self

    def numpy(self):
        return self.arr

    def item(self):
        return self.arr.reshape(-1)[0].item()


def empty(shape, dtype=np.uint8, device=None):
    if isinstance(shape, int):
        shape = (min(shape, 1 << 20),)  # the L2-flush buffer does not need to be 512 MiB here
    return _Tensor(np.zeros(shape, dtype))


def tensor(data, dtype=np.float64, device=None):
    return _Tensor(np.asarray(data, dtype))


def device(kind, index=0):
    return (kind, index)


class _Event:
    def __init__(self, enable_timing=False):
        self.t = 0.0

    def record(self, stream=None):
        self.t = time.perf_counter()

    def elapsed_time(self, other):
        return (other.t - self.t) * 1e3


class cuda:
    Event = _Event

    @staticmethod
    def is_available():
        return True

    @staticmethod
    def set_device(i):
        pass

    @staticmethod
    def synchronize():
        pass

    @staticmethod
    def ExternalStream(ptr, device=None):
        return ptr
