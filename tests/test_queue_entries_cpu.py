"""Index arithmetic of the 16-bit survivor-queue entries and of the swizzled / padded list-1 tiles
(hbt_kernels_v3.cuh: V3Smem::SWZ, v3_run_unit; hbt_kernels_v4.cuh: V4Smem::TIP), restated in numpy: every particle of a
sub-tile gets its own slot, the entry built in the prefilter decodes to that slot and to the list-2 position, and the
particles a drain round gathers for neighbouring lanes fall into different shared-memory banks."""
import numpy as np


def v3_slot(lane, s, sub):
    swz = 16 // (sub // 32)
    return (lane + 32 * s) ^ (s * swz)


def v3_entry(lane, s, j, sub):
    swz = 16 // (sub // 32)
    ej = (lane << 8) + j
    return ((ej ^ ((s * swz) << 8)) + (s << 13)) & 0xFFFF


def test_v3_slots_are_a_permutation_and_entries_decode():
    for sub, tj in ((64, 64), (128, 128)):
        ipl = sub // 32
        slots = {v3_slot(l, s, sub) for l in range(32) for s in range(ipl)}
        assert slots == set(range(sub))
        for l in range(32):
            for s in range(ipl):
                for j in (0, 1, tj // 2, tj - 1):
                    e = v3_entry(l, s, j, sub)
                    assert e >> 8 == v3_slot(l, s, sub) and e & 0xFF == j


def test_v3_neighbouring_lanes_hit_different_banks():
    """8-byte elements: 16 per 128-byte row.  The s-th particles of up to 16 / IPL neighbouring lanes must not share a
    bank pair (a half-warp of a drain round reads the particles of a few neighbouring lanes)."""
    for sub in (64, 128):
        ipl = sub // 32
        n = 16 // ipl
        for l0 in range(0, 32 - n + 1):
            banks = [v3_slot(l, s, sub) % 16 for l in range(l0, l0 + n) for s in range(ipl)]
            if l0 % n == 0:  # aligned groups of lanes: all distinct
                assert len(set(banks)) == len(banks), (sub, l0, banks)


def v4_slot(lane, s):
    return lane + 34 * s


def test_v4_padded_slots_and_entries():
    slots = [v4_slot(l, s) for l in range(32) for s in range(4)]
    assert len(set(slots)) == 128 and max(slots) < 128 + 2 * 3
    for l in range(32):
        for s in range(4):
            for j in (0, 77, 127):
                e = ((l << 7) + j + s * (34 << 7)) & 0xFFFF  # V4Smem::ESH = 7: nine bits of slot, seven of position
                sl = e >> 7
                assert sl == v4_slot(l, s) and e & 0x7F == j
                assert sl - 2 * (sl // 34) == l + 32 * s  # slot -> particle (parked pairs)
    # 16-byte records, 8 per 128-byte row: the four particles of two neighbouring lanes in eight different bank groups
    for l0 in range(0, 31, 2):
        groups = [v4_slot(l, s) % 8 for l in (l0, l0 + 1) for s in range(4)]
        assert len(set(groups)) == 8
