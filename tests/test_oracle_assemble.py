"""Oracle of the ranging aggregation (PosGenerator, Posgenerator.cpp:143-281) against hand-computed
cases (the pinning against the reference's own PosGenerator is in test_oracle_golden.py and
test_oracle_vs_ref.py)."""
import numpy as np

from oracle import oracle_py as O


def run(msgs, M=4, T=8, **kw):
    a, s, r, t = (np.array(c) for c in zip(*[(m[0], m[1], m[2], m[3]) for m in msgs]))
    err = np.array([m[4] if len(m) > 4 else 0.0 for m in msgs]) if any(len(m) > 4 for m in msgs) else None
    out = O.assemble(a[:, None], s[:, None], r[:, None], t[:, None], M, T, err=None if err is None else err[:, None], **kw)
    n = int(out["n_epochs"][0])
    return n, out["ranges"][:, :, 0], out["dt"][:, 0], out["err"][:, :, 0]


def test_epoch_closes_on_next_seq_and_on_the_timer():
    # seq 7: anchors 0,1,2 within 10 ms; seq 8 starts at t=0.1 -> report 1 was already sent by the
    # timer at 0.012+0.05 (gap 0.088 > 0.05) AND is sent again when seq 8 arrives (as written);
    # seq 8 is closed by the timer after the last ranging
    msgs = [(0, 7, 1000, 0.010), (1, 7, 2000, 0.011), (2, 7, 3000, 0.012), (0, 8, 1100, 0.100), (3, 8, 4100, 0.101)]
    n, r, dt, _ = run(msgs)
    assert n == 3
    assert r[0].tolist() == [1000, 2000, 3000, -1] and r[1].tolist() == [1000, 2000, 3000, -1]
    assert r[2].tolist() == [1100, -1, -1, 4100]
    assert dt[0] == 0.1 and np.isclose(dt[1], 0.100 - 0.062) and np.isclose(dt[2], 0.151 - 0.100)
    assert (dt[3:] == -1).all() and (r[3:] == -1).all()


def test_dense_stream_no_timer():
    # 10 Hz epochs but rangings every 20 ms: the timer never fires, one report per sequence number
    msgs = [(a, s, 1000 * s + a, 0.1 * s + 0.02 * a) for s in range(5) for a in range(4)]
    n, r, dt, _ = run(msgs)
    assert n == 5  # 4 closed by the next seq + the last one by the timer
    for s in range(5):
        assert r[s].tolist() == [1000 * s + a for a in range(4)]
    assert dt[0] == 0.1 and np.allclose(dt[1:4], 0.1) and np.isclose(dt[4], 0.46 + 0.05 - 0.4)


def test_later_ranging_of_the_same_anchor_overwrites_and_error_rule():
    # within a sequence errorEstimation only overwrites when > 0 (Posgenerator.cpp:94,236); the first
    # ranging of a sequence stores it unconditionally (:266)
    msgs = [(1, 3, 500, 0.00, 0.0), (0, 3, 700, 0.01, 0.02), (0, 3, 720, 0.02, 0.0), (1, 3, 510, 0.03, 0.05)]
    n, r, dt, e = run(msgs)
    assert n == 1 and r[0].tolist() == [720, 510, -1, -1]
    assert e[0].tolist() == [0.02, 0.05, 0.0, 0.0]


def test_b12_stale_slots_survive_a_wrap_of_the_sequence_number():
    # seq 5 with anchors 0..3, then 255 other sequence numbers with anchor 0 only, then seq 5 again with
    # anchor 0: as written only slot 0 of the row is cleared, so anchors 1..3 of 256 epochs ago leak in
    msgs = [(a, 5, 9000 + a, 0.001 * a) for a in range(4)]
    t = 0.1
    for k in range(1, 256):
        msgs.append((0, (5 + k) % 256, 100 + k, t)); t += 0.02
    msgs.append((0, 5, 4242, t))
    n, r, dt, _ = run(msgs, T=300)
    # reports: seq 5 by the timer (gap 0.097 s), then one per new sequence number (256 arrivals), then
    # the wrapped seq 5 by the final timer
    assert n == 258
    assert r[0].tolist() == r[1].tolist() == [9000, 9001, 9002, 9003]
    assert r[257].tolist() == [4242, 9001, 9002, 9003]
    n2, r2, _, _ = run(msgs, T=300, fix_b12=True)
    assert n2 == 258 and r2[257].tolist() == [4242, -1, -1, -1]
    assert np.array_equal(r[:257], r2[:257])


def test_padding_and_truncation():
    msgs = [(0xFF, 0, 0, 0.0), (0, 1, 10, 0.5), (9, 1, 11, 0.51), (1, 1, 12, 0.52), (0, 2, 20, 0.6), (0, 3, 30, 0.7)]
    n, r, dt, _ = run(msgs, T=2)
    # gaps of 80 / 100 ms: every sequence is reported by the timer and again when the next one starts
    assert n == 5  # five reports exist, two fit; anchor 9 >= M and 0xFF are padding
    assert r[0].tolist() == [10, 12, -1, -1] and r[1].tolist() == [10, 12, -1, -1]
    assert np.isclose(dt[1], 0.6 - 0.57)


def test_empty_log():
    out = O.assemble(np.full((5, 3), 0xFF), np.zeros((5, 3)), np.zeros((5, 3)), np.zeros((5, 3)), 4, 3)
    assert (out["n_epochs"] == 0).all() and (out["dt"] == -1).all() and (out["ranges"] == -1).all()
