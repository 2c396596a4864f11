"""GPU: the device-side Monte Carlo input generator (kfpos_synth_k8) and the chunked K8 run on it."""
import numpy as np
import pytest

from roskfpos_b200 import synth

pytestmark = pytest.mark.gpu


def test_generator_is_a_pure_function_of_seed_filter_and_event(kflib):
    """Same (seed, global filter, global event) -> same value, however the batch is sharded over GPUs
    or cut into chunks; different seeds differ."""
    anc = synth.anchors_for(8)
    whole = synth.k8_montecarlo_chunk(4096, 0, 4, anc, "cuda:0", seed=5, full=True, want_x0=True)
    # shard: filters [1000, 1500) generated on their own
    part = synth.k8_montecarlo_chunk(500, 0, 4, anc, "cuda:0", seed=5, full=True, first_filter=1000, want_x0=True)
    for k in ("ranges", "sensors", "x0", "truth_end"):
        assert np.array_equal(whole[k][..., 1000:1500].cpu().numpy(), part[k].cpu().numpy()), k
    # chunk: macro-steps 2..3 generated on their own
    tail = synth.k8_montecarlo_chunk(4096, 2, 2, anc, "cuda:0", seed=5, full=True)
    assert np.array_equal(whole["ranges"][2:].cpu().numpy(), tail["ranges"].cpu().numpy())
    ns = whole["sensors"].shape[0] // 2
    assert np.array_equal(whole["sensors"][ns:].cpu().numpy(), tail["sensors"].cpu().numpy())
    assert np.array_equal(whole["truth_end"].cpu().numpy(), tail["truth_end"].cpu().numpy())
    other = synth.k8_montecarlo_chunk(4096, 0, 4, anc, "cuda:0", seed=6, full=True)
    assert not np.array_equal(whole["ranges"].cpu().numpy(), other["ranges"].cpu().numpy())


def test_generator_statistics(kflib):
    """Noise levels of the samples around the truth they are drawn for (config_imu.xml / config_mag.xml
    values, 0.10 m ranging noise), zero mean, no correlation between filters or events."""
    anc = synth.anchors_for(8)
    N = 200000
    w = synth.k8_montecarlo_chunk(N, 0, 2, anc, "cuda:0", seed=9, full=False, want_x0=True)
    x0 = w["x0"].cpu().numpy()
    sens = w["sensors"].cpu().numpy()
    rng = w["ranges"].cpu().numpy()
    # reconstruct the truth from x0: phases from position / velocity at t = 0
    pa = np.arctan2((x0[0] - 5) / 3, x0[2] / 0.6); pb = np.arctan2((x0[1] - 5) / 3, x0[3] / 0.93)
    th0 = x0[6]
    t, row, k_toa = 0.0, 0, 0
    gyro, acc, comp, rres = [], [], [], []
    for kind, dt, off, aux in w["events"]:
        t += dt
        px = 5 + 3 * np.sin(0.20 * t + pa); py = 5 + 3 * np.sin(0.31 * t + pb)
        ax = -0.12 * np.sin(0.20 * t + pa); ay = -0.2883 * np.sin(0.31 * t + pb)
        th = th0 + 0.05 * t
        if kind == synth.EV_IMU:
            gyro.append(sens[off] - 0.05)
            acc.append(sens[off + 1] - (np.cos(th) * ax + np.sin(th) * ay))
            acc.append(sens[off + 2] - (-np.sin(th) * ax + np.cos(th) * ay))
        elif kind == synth.EV_COMPASS:
            d = sens[off] - th
            comp.append(d - 2 * np.pi * np.round(d / (2 * np.pi)))
        else:
            d = np.sqrt((px[None] - anc[:, 0:1]) ** 2 + (py[None] - anc[:, 1:2]) ** 2 + (1.049 - anc[:, 2:3]) ** 2)
            rres.append(rng[k_toa] / 1000.0 + 0.0005 - d)  # + half a millimetre: the floor
            k_toa += 1
    for name, v, var in (("gyro", np.array(gyro), 0.089), ("acc", np.array(acc), 0.003), ("compass", np.array(comp), 1e-4),
                         ("range", np.array(rres), 0.01)):
        assert abs(v.mean()) < 4 * np.sqrt(var / v.size) + 1e-6, (name, v.mean())
        assert abs(v.var() / var - 1) < 0.01, (name, v.var())
    g = np.array(gyro)
    assert abs(np.corrcoef(g[0], g[1])[0, 1]) < 0.01           # successive events of one filter
    assert abs(np.corrcoef(g[0][:-1], g[0][1:])[0, 1]) < 0.01  # neighbouring filters
    a = np.array(acc)
    assert abs(np.corrcoef(a[0], a[1])[0, 1]) < 0.01           # the two samples of one Philox block pair


def test_chunked_monte_carlo_equals_single_shot_and_the_oracle(kflib, oracle):
    """A K8 Monte Carlo run generated and replayed chunk by chunk (what a 64 M-filter run does) gives
    bit-identical states to the single-shot run, and a sample of its filters agrees with the oracle
    replaying the same generated inputs."""
    from roskfpos_b200.batch import Batch
    anc = synth.anchors_for(8)
    N, n_macro = 20000, 6
    one = synth.k8_montecarlo_chunk(N, 0, n_macro, anc, "cuda:0", seed=21, full=True, want_x0=True)
    with Batch(kflib.MODEL_K8, N, anchors=anc, xml=synth.K8_XML, accel_noise=0.5, jolt=0.5) as b:
        b.set_state(one["x0"])
        b.replay_events(one["events"], ranges=one["ranges"], sensors=one["sensors"], err=0.01)
        x1, P1, st1 = b.get_state()
    with Batch(kflib.MODEL_K8, N, anchors=anc, xml=synth.K8_XML, accel_noise=0.5, jolt=0.5) as b:
        b.set_state(one["x0"])
        bufs = None
        for m0 in range(0, n_macro, 2):
            bufs = synth.k8_montecarlo_chunk(N, m0, 2, anc, "cuda:0", seed=21, full=True, out=bufs)
            b.replay_events(bufs["events"], ranges=bufs["ranges"], sensors=bufs["sensors"], err=0.01)
        x2, P2, st2 = b.get_state()
        stats = b.error_stats(bufs["truth_end"])
    assert np.array_equal(x1, x2) and np.array_equal(P1, P2) and np.array_equal(st1, st2)
    assert np.sqrt(stats[1] / stats[2]) < 0.05  # tracking: RMSE in x,y of a few centimetres
    idx = np.arange(0, N, 97)
    cfg = oracle.k8_cfg(0.5, 0.5, **synth.K8_ORACLE_CFG)
    ref = oracle.k8_replay(one["x0"][:, idx].cpu().numpy(), None, one["events"], one["ranges"][:, :, idx].cpu().numpy(),
                           one["sensors"][:, idx].cpu().numpy(), anc, 0.01, cfg)
    assert np.abs(x1[:, idx] - ref["x"]).max() < 1e-9
    assert (np.abs(P1[:, idx] - ref["P"]).max(axis=0) / np.abs(ref["P"]).max(axis=0)).max() < 1e-9
