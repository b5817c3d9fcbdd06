"""bench.py's synthetic decoder keeps its own DPB: the slot pools must never hand out a slot whose content is still
referenced, and the assignment must repeat after STEP_VARIANTS intra periods (the prepared pictures are replayed)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def test_dpb_slot_policy_has_no_hazards_and_is_periodic():
    seq, key_slot, _ = bench.build_workload(64, 64, 8)
    assert len(seq) == 32 * bench.STEP_VARIANTS
    holder = {key_slot: 0}  # slot -> POC it holds (POC 0 = the uploaded reference)
    poc_base = 0
    for rep in range(3):  # replayed: the state at the end of the last variant is the state the first variant expects
        for p in seq:
            poc = poc_base + p.params.poc
            g, off = divmod(p.params.poc - 1, 8)
            off += 1
            r0, r1 = bench.GOP_REFS[off]
            want = {poc_base + g * 8 + r0} | ({poc_base + g * 8 + r1} if r1 is not None else set())
            refs = {int(s) for s in p.pus["ref_slot"].ravel() if s >= 0} if len(p.pus) else set()
            assert {holder[s] for s in refs} <= want, f"POC {poc}: a reference slot holds the wrong picture"
            d = p.params.dst_slot
            assert d not in refs
            # the picture being overwritten must not be needed any more: nothing newer than 16 pictures old is ever referenced
            assert d not in holder or poc - holder[d] >= 16, f"POC {poc} overwrites POC {holder[d]} too early"
            holder[d] = poc
        poc_base += 32 * bench.STEP_VARIANTS


def test_tight_dpb_variant_still_references_the_right_pictures(monkeypatch):
    """B200_TIGHT_DPB=1 (the slot-renaming experiment of tools/sweep_bench.py): 7 slots, every slot reused as early as the GOP
    structure allows — references must still hold the pictures the GOP names, and the assignment must stay periodic."""
    import importlib
    monkeypatch.setenv("B200_TIGHT_DPB", "1")
    tight = importlib.reload(bench)
    try:
        assert (tight.KEY_SLOTS, tight.REFB_SLOTS, tight.NONREF_SLOTS) == (2, 3, 2)
        seq, key_slot, _ = tight.build_workload(64, 64, 8)
        holder = {key_slot: 0}
        poc_base = 0
        for rep in range(3):
            for p in seq:
                poc = poc_base + p.params.poc
                g, off = divmod(p.params.poc - 1, 8)
                off += 1
                r0, r1 = tight.GOP_REFS[off]
                want = {poc_base + g * 8 + r0} | ({poc_base + g * 8 + r1} if r1 is not None else set())
                refs = {int(s) for s in p.pus["ref_slot"].ravel() if s >= 0} if len(p.pus) else set()
                assert {holder[s] for s in refs} <= want, f"POC {poc}: a reference slot holds the wrong picture"
                assert p.params.dst_slot not in refs
                holder[p.params.dst_slot] = poc
            poc_base += 32 * tight.STEP_VARIANTS
    finally:
        monkeypatch.delenv("B200_TIGHT_DPB")
        importlib.reload(bench)
