"""Exact-set stand-in for the third-party `pybloom_live` package (not installed, unpinned in the
reference).  TEST INFRASTRUCTURE ONLY: lets /root/reference/Code/utils.py:9 import so the unmodified
reference Modules.py / main.py functions can be executed when generating golden vectors."""


class BloomFilter:
    def __init__(self, capacity=10, error_rate=1e-3):
        self.capacity = capacity
        self.error_rate = error_rate
        self._s = set()

    def add(self, key):
        self._s.add(tuple(int(v) for v in key))

    def __contains__(self, key):
        return tuple(int(v) for v in key) in self._s

    def __len__(self):
        return len(self._s)
