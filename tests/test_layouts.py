"""Boundary fixtures the reference pins in its own tests: struct sizes, key
encodings and the arithmetic of the control-plane helpers, checked against the Python layouts
and synthetic generators used by tests and bench.py."""
import numpy as np

from bng_b200 import layouts as L
from bng_b200 import synth as S


def test_struct_sizes_match_the_c_layouts():
    # test/ebpf/maps_test.go:68-130 and gcc-verified sizes of SURVEY.md §8a
    sizes = {"subscriber_binding": 24, "antispoof_config": 8, "antispoof_stats": 48, "spoof_event": 56, "token_bucket": 32,
             "qos_stats": 32, "nat_key": 16, "eim_key": 8, "eim_mapping": 32, "nat_session": 80, "port_block": 32,
             "subscriber_nat": 64, "nat_pool_entry": 16, "nat_stats": 104, "nat_log_entry": 40, "alg_config": 8,
             "nat_config": 16, "pool_assignment": 25, "vlan_key": 4, "ip_pool": 28, "dhcp_server_config": 16,
             "dhcp_stats": 80, "circuit_id_key": 32}
    for name, sz in sizes.items():
        assert getattr(L, name).itemsize == sz, name
    assert L.nat_session.fields["last_seen"][1] == 24 and L.nat_session.fields["state"][1] == 72
    assert L.port_block.fields["next_port"][1] == 8 and L.port_block.fields["subscriber_id"][1] == 24
    assert L.pool_assignment.fields["lease_expiry"][1] == 13


def test_mac_and_ip_encodings():
    # pkg/ebpf/loader_test.go:16-130: aa:bb:cc:dd:ee:ff -> 0xAABBCCDDEEFF, 10.0.1.1 -> 0x0A000101
    assert S.mac_bytes(np.uint64(0xAABBCCDDEEFF)).tolist() == [0xAA, 0xBB, 0xCC, 0xDD, 0xEE, 0xFF]
    assert S.ip_bytes(np.uint32(0x0A000101)).tolist() == [10, 0, 1, 1]
    assert int(S.sub_mac_key(5)) == 0x020000000005 and int(S.sub_ip(5)) == 0x64400005


def test_port_block_rule_of_allocate_nat():
    # pkg/nat/manager_test.go:164-247: 63 subscribers per public IP, blocks of 1024 from 1024
    k, v, pubs = S.nat_blocks(130)
    assert v["block"]["port_start"][0] == 1024 and v["block"]["port_end"][0] == 2047
    assert v["block"]["port_start"][62] == 1024 + 62 * 1024 and v["block"]["port_end"][62] == 65535
    assert v["block"]["port_start"][63] == 1024  # 64th subscriber starts the next public IP
    assert len(pubs) == 3 and (v["block"]["next_port"] == v["block"]["port_start"]).all()
    assert v["block"]["subscriber_id"].tolist() == list(range(1, 131))
    assert (v["block"]["block_size_log2"] == 10).all()


def test_qos_burst_defaults():
    # pkg/qos/manager.go:181-209
    assert S.qos_default_burst(100_000_000) == 10 * 1024 * 1024
    assert S.qos_default_burst(20_000_000) == 2_500_000
    assert S.qos_default_burst(100_000) == 65536
    k, v = S.qos_buckets(8, upload=True)
    assert v["burst_bytes"][1] == 2_500_000 and v["rate_bps"][1] == 20_000_000 and (v["tokens"] == v["burst_bytes"]).all()
    k, v = S.qos_buckets(8, upload=False)
    assert v["burst_bytes"][1] == 2_000_000 and v["burst_bytes"][7] == 65536  # policy burst, default rule for 'unlimited'


def test_imix_mix():
    l = S.imix_lengths(120000, 1)
    assert abs(l.mean() - 361.83) < 3 and set(np.unique(l)) == {64, 594, 1518}
