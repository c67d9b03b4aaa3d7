"""Property tests (hypothesis) of the host logic behind the sliding-window schedule and the two multi-GPU partitions
(segmantic_b200/seg/sliding_window.py): random volume sizes, roi sizes, overlaps and world sizes.

* schedule == oracle (MONAI dense_patch_slices restatement), every voxel covered, windows inside the padded volume;
* slab partition: contiguous output slabs, every rank executes exactly the window rows that intersect its slab;
* window ownership: the ranks' window ranges tile the window list without gaps or repeats, their output planes tile axis 0,
  every rank's blend rows are covered by its own + the received windows, and send / receive ranges match."""
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import sliding_window as osw
from segmantic_b200.seg import sliding_window as sw

dims = st.tuples(st.integers(1, 400), st.integers(1, 200), st.integers(1, 200))
rois = st.tuples(st.integers(1, 128), st.integers(1, 96), st.integers(1, 96))
overlaps = st.sampled_from([0.0, 0.25, 0.5, 0.75, 0.9])


@settings(max_examples=120, deadline=None)
@given(size=dims, roi=rois, overlap=overlaps)
def test_schedule_matches_oracle_and_covers(size, roi, overlap):
    s = sw.make_schedule(size, roi, overlap, "constant")
    assert s.windows() == [tuple(w) for w in osw.window_starts(s.padded_size, roi, overlap)]
    for a in range(3):
        starts, n = s.starts[a], s.padded_size[a]
        assert starts == sorted(set(starts)) and starts[0] == 0 and starts[-1] + roi[a] == n
        assert all(b - a_ <= roi[a] for a_, b in zip(starts, starts[1:]))     # no uncovered gap between windows


@settings(max_examples=120, deadline=None)
@given(size=dims, roi=rois, overlap=overlaps, world=st.integers(1, 8))
def test_slab_partition_properties(size, roi, overlap, world):
    s = sw.make_schedule(size, roi, overlap, "constant")
    parts = sw.slab_partition(s, world)
    assert len(parts) == world and parts[0]["x0"] == 0 and parts[-1]["x1"] == s.padded_size[0]
    s0 = s.starts[0]
    for a, b in zip(parts, parts[1:]):
        assert a["x1"] == b["x0"] and a["x0"] <= a["x1"]
    for p in parts:
        rows = [j for j in range(len(s0)) if p["x1"] > p["x0"] and s0[j] < p["x1"] and s0[j] + roi[0] > p["x0"]]
        assert list(range(p["a0_begin"], p["a0_end"])) == rows
        if rows:
            assert p["vol_x0"] == s0[rows[0]] <= p["x0"] and p["vol_x1"] == s0[rows[-1]] + roi[0] >= p["x1"]


@settings(max_examples=120, deadline=None)
@given(size=dims, roi=rois, overlap=overlaps, world=st.integers(1, 8))
def test_window_partition_properties(size, roi, overlap, world):
    s = sw.make_schedule(size, roi, overlap, "constant")
    try:
        parts = sw.window_partition(s, world)
    except ValueError:
        return  # documented: too many ranks for the number of window rows
    per_row = len(s.starts[1]) * len(s.starts[2])
    total = len(s.starts[0]) * per_row
    s0, size0 = s.starts[0], s.padded_size[0]
    assert parts[0]["w_lo"] == 0 and parts[-1]["w_hi"] == total
    for a, b in zip(parts, parts[1:]):
        assert a["w_hi"] == b["w_lo"]                                           # the window list is tiled
    live = [p for p in parts if p["w_hi"] > p["w_lo"]]
    assert live[0]["x0"] == 0 and live[-1]["x1"] == size0
    for a, b in zip(live, live[1:]):
        assert a["x1"] == b["x0"]                                               # output planes are tiled
        assert a["send_lo"] == max(b["wb"], a["w_lo"]) and b["wb"] >= a["w_lo"]  # what a sends is what b misses
    for p in live:
        if p["x1"] <= p["x0"]:
            continue
        rows = [j for j in range(len(s0)) if s0[j] < p["x1"] and s0[j] + roi[0] > p["x0"]]
        assert (p["b_begin"], p["b_end"]) == (rows[0], rows[-1] + 1)
        # every window of the rows it blends is either its own or received from the previous rank
        assert p["wb"] <= rows[0] * per_row and (rows[-1] + 1) * per_row <= p["w_hi"]
        assert p["vol_x0"] <= s0[p["w_lo"] // per_row] and p["vol_x1"] >= s0[(p["w_hi"] - 1) // per_row] + roi[0]
