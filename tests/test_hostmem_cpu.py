"""CPU checks of the staging-memory placement helper (strotss_tensorflow_b200/hostmem.py): the sysfs cpulist parser, and that
placement degrades to a recorded no-op -- affinity and memory policy restored -- where the GPU's NUMA node is unknown."""
import os

from strotss_tensorflow_b200 import hostmem


def test_cpulist_parser():
    assert hostmem._parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert hostmem._parse_cpulist("") == set()
    assert hostmem._parse_cpulist("5") == {5}


def test_unknown_gpu_gives_empty_record(monkeypatch):
    monkeypatch.setattr(hostmem, "_bus_id", lambda index: None)
    assert hostmem.gpu_numa(0) == {"bus_id": None, "numa_node": None, "local_cpus": None}
    monkeypatch.setattr(hostmem, "_bus_id", lambda index: "ffff:ff:1f.0")           # no such device in sysfs
    info = hostmem.gpu_numa(0)
    assert info["bus_id"] == "ffff:ff:1f.0" and info["numa_node"] is None and info["local_cpus"] is None


def test_numa_local_is_a_recorded_noop_without_topology(monkeypatch):
    monkeypatch.setattr(hostmem, "_bus_id", lambda index: None)
    before = os.sched_getaffinity(0)
    rec = {}
    with hostmem.numa_local(3, rec) as r:
        assert r is rec and os.sched_getaffinity(0) == before
    assert rec["gpu"] == 3 and rec["cpus_bound"] is None and rec["mempolicy"] == "unchanged" and rec["allowed_cpus"] == len(before)
    assert os.sched_getaffinity(0) == before


def test_numa_local_binds_and_restores(monkeypatch):
    allowed = os.sched_getaffinity(0)
    one = min(allowed)
    monkeypatch.setattr(hostmem, "gpu_numa", lambda index: {"bus_id": "0000:00:00.0", "numa_node": -1, "local_cpus": [one, 10 ** 6]})
    rec = {}
    with hostmem.numa_local(0, rec):
        assert os.sched_getaffinity(0) == {one}
    assert os.sched_getaffinity(0) == allowed
    assert rec["cpus_bound"].startswith(f"{one}-{one} (1 of the 2 CPUs")
    assert rec["mempolicy"] == "unchanged"                                            # node -1: no policy call
    # a GPU whose local CPUs are all outside this process's cpuset: nothing is bound, and the record says why
    monkeypatch.setattr(hostmem, "gpu_numa", lambda index: {"bus_id": "0000:00:00.0", "numa_node": -1, "local_cpus": [10 ** 6]})
    rec = {}
    with hostmem.numa_local(0, rec):
        assert os.sched_getaffinity(0) == allowed
    assert "none of the GPU's local CPUs" in rec["cpus_bound"]
