"""ORACLE (test infrastructure, NOT product code): plain-Python restatement of the reference's ReplayBuffer ring
(plen_ros/src/plen_ros_helpers/td3.py:122-193): a list that grows to max_size, then is overwritten from index 0 upwards.
Pinned by tests/golden/td3_golden.npz (ring order produced by the reference class itself).  Only tests/ may import this."""


class ReplayRing:
    def __init__(self, max_size):
        self.storage, self.max_size, self.ptr = [], int(max_size), 0

    def add(self, data):                       # td3.py:136-147
        if len(self.storage) == self.max_size:
            self.storage[int(self.ptr)] = data
            self.ptr = (self.ptr + 1) % self.max_size
        else:
            self.storage.append(data)
