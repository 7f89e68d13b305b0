"""Parity tests proper: the CUDA path, called through the C ABI of libft8b200.so, against the CPU oracle
on the same seeded inputs, against the committed golden fixtures (made by the unmodified reference), and
-- at BASELINE.json's full sizes -- through size-independent properties.

Bars (BASELINE.json north_star): candidate lists, LDPC hard decisions, CRC and message text bit-exact;
decimator floats within 1e-5 relative (asserted BIT-EXACT here, which is stronger); waterfall cells within
+-1 LSB on <= 0.01 % of cells (asserted IDENTICAL here)."""
import numpy as np
import pytest

from conftest import bits_equal, golden
from oracle.pyoracle import cand_dtype, msg_dtype, result_dtype, status_dtype
from tools import ft8enc, synth

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def dev():
    return torch.device("cuda:0")


def view(t, dt):
    a = t.cpu().numpy()
    return a.view(dt).reshape(a.shape[:-1])


def make_slot(oracle, sigs, seed, condition=True):
    i_s, q_s = synth.slot_f32(sigs, seed)
    if condition:
        i_s, q_s, _ = oracle.condition(i_s, q_s, 48000)
    return i_s, q_s


def std_sig(f0=700.0, t0=0.5, snr=-10.0, msg=("CQ", "K1JT", "FN20")):
    return (ft8enc.tones(ft8enc.pack_std(*msg)), f0, t0, snr)


# ------------------------------------------------------------------------------------------- decimator
def oracle_decim(oracle, iq):
    n8 = (iq.size // 8) * 8
    return oracle.decimate_slot(iq[:n8], want_y2=True)


@pytest.mark.parametrize("nbytes", [12016 * 40, 12016 * 3 + 1502 * 4 + 8 * 13, 1502 * 12, 1496, 8])
def test_decimator_random_streams(ctx, oracle, nbytes):
    """uint8 IQ incl. 0x00/0xFF (the int8 negate wrap), full super-blocks, ragged tails, < 1 block."""
    rng = np.random.default_rng(nbytes)
    stride = (nbytes + 15) // 16 * 16
    iq = rng.integers(0, 256, size=(3, stride), dtype=np.uint8)
    iq[0, ::53] = 0
    iq[1, 7::97] = 255
    iq[2] = rng.integers(118, 139, size=stride, dtype=np.uint8)
    d_i, d_q, cnt, peak, y2 = ctx.decimate(torch.from_numpy(iq).to(dev()), 3, nbytes, stride=stride, want_y2=True)
    torch.cuda.synchronize()
    for s in range(3):
        oi, oq, oy2i, oy2q = oracle_decim(oracle, iq[s, :nbytes])
        n = int(cnt[s])
        assert n == oi.size == (nbytes // 2) // 751
        gi, gq, gy = d_i[s].cpu().numpy(), d_q[s].cpu().numpy(), y2[s].cpu().numpy()
        assert np.array_equal(gy[:n, 0], oy2i) and np.array_equal(gy[:n, 1], oy2q), "integer CIC core must be exact"
        assert bits_equal(gi[:n], oi) and bits_equal(gq[:n], oq), "float outputs bit-exact (spec: <= 1e-5 relative)"
        assert not gi[n:].any() and not gq[n:].any(), "tail of the 48000-sample buffer is zero (decoder(), rtlsdr_ft8d.c:243-246)"
        expect_peak = max(np.abs(oi).max(initial=0), np.abs(oq).max(initial=0))
        assert float(peak[s]) == float(expect_peak)


@pytest.mark.parametrize("fill", [0x00, 0xFF, 0x80, 0x7F, 0x01])
def test_decimator_constant_input(ctx, oracle, fill):
    nbytes = 12016 * 12
    iq = np.full(nbytes, fill, np.uint8)
    d_i, d_q, cnt, peak, y2 = ctx.decimate(torch.from_numpy(iq).to(dev()), 1, nbytes, want_y2=True)
    oi, oq, oy2i, oy2q = oracle_decim(oracle, iq)
    n = int(cnt[0])
    assert np.array_equal(y2[0, :n, 0].cpu().numpy(), oy2i) and np.array_equal(y2[0, :n, 1].cpu().numpy(), oy2q)
    assert bits_equal(d_i[0, :n].cpu().numpy(), oi) and bits_equal(d_q[0, :n].cpu().numpy(), oq)
    if fill == 0x80:
        assert not oi.any() and float(peak[0]) == 0.0


def test_decimator_golden(ctx):
    g = golden("decim_random")
    iq = g["iq"]
    nbytes = (iq.size // 8) * 8
    buf = np.zeros((nbytes + 15) // 16 * 16, np.uint8)
    buf[:nbytes] = iq[:nbytes]
    d_i, d_q, cnt, _, _ = ctx.decimate(torch.from_numpy(buf).to(dev()), 1, nbytes, stride=buf.size)
    n = int(cnt[0])
    assert n == g["i"].size
    assert bits_equal(d_i[0, :n].cpu().numpy(), g["i"]) and bits_equal(d_q[0, :n].cpu().numpy(), g["q"])


@pytest.fixture(scope="module")
def raw_slot():
    """BASELINE config #2: one 15 s slot of raw 2.4 Msps uint8 IQ with one message, 36 M complex samples."""
    return synth.raw_u8([std_sig(800.0, 0.5, 20.0)], 3)


def test_full_raw_slot(ctx, oracle, raw_slot):
    d_raw = torch.from_numpy(raw_slot).to(dev())
    d_i, d_q, cnt, peak, _ = ctx.decimate(d_raw, 1, raw_slot.size)
    oi, oq = oracle.decimate_slot(raw_slot)
    assert int(cnt[0]) == oi.size == 47936
    ri = np.zeros(48000, np.float32); rq = np.zeros(48000, np.float32)
    ri[:oi.size] = oi; rq[:oq.size] = oq
    assert bits_equal(d_i[0].cpu().numpy(), ri) and bits_equal(d_q[0].cpu().numpy(), rq)
    # whole path: raw -> spots, against oracle decimate + decoder() conditioning + ft8_subsystem
    ic, qc, _ = oracle.condition(ri, rq, oi.size)
    o = oracle.subsystem(ic, qc)
    ctx.process_raw(d_raw, 1)
    res, n = ctx.fetch_results(1)
    assert n[0] == o["n"] >= 1
    assert res[0].tobytes() == o["results"].tobytes()
    assert res[0][0]["call"] == b"K1JT" and res[0][0]["loc"] == b"FN20" and res[0][0]["freq"] == 800
    # end-to-end host entry point gives the same bytes
    res2, n2 = ctx.process_raw_host(raw_slot, 1)
    assert n2[0] == n[0] and res2.tobytes() == res.tobytes()


def test_raw_batch_properties(ctx, raw_slot):
    """Full-size properties: every copy of a slot in a batch decodes identically (position independence),
    a silent slot (all 0x80) yields zero samples and no candidates, and the integer CIC core is linear:
    y2(a) + y2(b) == y2(a + b - 128) when no byte saturates or hits the -128 wrap."""
    B = 4
    big = torch.from_numpy(raw_slot).to(dev()).repeat(B, 1).contiguous()
    big[2] = 0x80
    ctx.process_raw(big, B)
    res, n = ctx.fetch_results(B)
    assert n[0] == n[1] == n[3] >= 1 and n[2] == 0
    assert res[0].tobytes() == res[1].tobytes() == res[3].tobytes()
    rng = np.random.default_rng(5)
    nbytes = 12016 * 64
    a = rng.integers(100, 130, size=nbytes, dtype=np.uint8)
    b = rng.integers(110, 150, size=nbytes, dtype=np.uint8)
    c = (a.astype(np.int32) + b.astype(np.int32) - 128).astype(np.uint8)
    y = ctx.decimate(torch.from_numpy(np.stack([a, b, c])).to(dev()), 3, nbytes, want_y2=True)[4].cpu().numpy().astype(np.int64)
    n_out = nbytes // 2 // 751
    assert np.array_equal(y[0, :n_out] + y[1, :n_out], y[2, :n_out])


# ------------------------------------------------------------------------------------------- waterfall
@pytest.fixture(scope="module")
def slots(oracle):
    out = [make_slot(oracle, [std_sig()], 7)]
    i_s, q_s, _ = synth.crowded_band(ft8enc, 60, 99)
    out.append(tuple(oracle.condition(i_s, q_s, 48000)[:2]))
    i_s, q_s, _ = synth.crowded_band(ft8enc, 25, 5, snr_lo=-18, snr_hi=0)
    out.append(tuple(oracle.condition(i_s, q_s, 48000)[:2]))
    out.append(make_slot(oracle, [], 11))  # noise only
    out.append((np.zeros(48000, np.float32), np.zeros(48000, np.float32)))  # silence: every cell clamps to 0
    out.append(make_slot(oracle, [std_sig(50.0, 0.0, 5.0), std_sig(1500.0, 2.3, 0.0, ("K1ABC", "W9XYZ", "-15"))], 13))  # band/time edges
    return out


def test_waterfall_parity(ctx, oracle, slots):
    hi = np.stack([s[0] for s in slots]); hq = np.stack([s[1] for s in slots])
    mag = ctx.waterfall(torch.from_numpy(hi).to(dev()), torch.from_numpy(hq).to(dev())).cpu().numpy()
    for s in range(len(slots)):
        ref = oracle.waterfall(hi[s], hq[s])
        ndiff = int((mag[s] != ref).sum())
        assert ndiff == 0, f"slot {s}: {ndiff} of 94208 cells differ (spec allows +-1 LSB on <= 9 cells; this build is exact)"


def test_quantiser_is_exact_over_all_floats(ctx, oracle):
    """The waterfall kernel's dB quantiser is straight-line code: log2 estimate, ONE load of the two step thresholds around it,
    two compares.  That is exact iff the estimate is never off by more than one step; swept on the device over every
    non-negative float bit pattern, +inf and all NaNs against a search over the 256 host-computed thresholds, which are
    themselves the reference's log10f expression (tables.cu; oracle.quantize_db pins a sample of them here)."""
    bad, corrected, far = ctx.selfcheck_quantiser()
    assert bad == 0 and far == 0
    assert 0 < corrected < 2 ** 31 // 50          # the +-1 correction exists and is rare
    thr = oracle.db_thresholds()
    for k in (1, 17, 128, 255):
        below = np.nextafter(np.float32(thr[k]), np.float32(0))
        assert oracle.quantize_db(float(thr[k])) == k and oracle.quantize_db(float(below)) == k - 1


def test_waterfall_applies_decoder_conditioning(ctx, oracle):
    """Unconditioned samples + slot peak: the kernel applies decoder()'s 0.5/max scale on load (rtlsdr_ft8d.c:248-263)."""
    i_s, q_s = synth.slot_f32([std_sig()], 21)
    i_s *= np.float32(3.7e-3); q_s *= np.float32(3.7e-3)
    ic, qc, _ = oracle.condition(i_s, q_s, 48000)
    peak = torch.tensor([max(np.abs(i_s).max(), np.abs(q_s).max())], dtype=torch.float32, device=dev())
    d_i = torch.from_numpy(i_s[None]).to(dev()); d_q = torch.from_numpy(q_s[None]).to(dev())
    mag = ctx.waterfall(d_i, d_q, peak).cpu().numpy()[0]
    assert np.array_equal(mag, oracle.waterfall(ic, qc))
    ctx.condition(d_i, d_q, peak)
    assert bits_equal(d_i[0].cpu().numpy(), ic) and bits_equal(d_q[0].cpu().numpy(), qc)


# ------------------------------------------------------------------------------------------- sync / decode / spots
def check_slot_against_oracle(oracle, c, mag_np, cand, ncand, ok, stage, status, msg, plain, llr, res, nres, umsg, ufreq, uscore, s, K, M):
    o = oracle.decode_waterfall(mag_np, max_cand=K, max_msgs=M)
    gc = view(cand[s], cand_dtype)[: int(ncand[s])]
    assert int(ncand[s]) == len(o["cands"])
    assert np.array_equal(gc, o["cands"]), "candidate list must match in content AND order"
    g_st, g_msg = view(status[s], status_dtype), view(msg[s], msg_dtype)
    for k, cd in enumerate(o["cands"]):
        d = oracle.decode(mag_np, cd)
        assert int(ok[s, k]) == d["ok"], (s, k)
        assert bits_equal(llr[s, k].cpu().numpy(), d["llr"]), f"slot {s} cand {k}: normalised LLRs"
        assert np.array_equal(plain[s, k].cpu().numpy(), d["plain"]), f"slot {s} cand {k}: LDPC hard decisions"
        assert g_st[k]["ldpc_errors"] == d["status"]["ldpc_errors"]
        st = int(stage[s, k])
        assert st == (1 if d["status"]["ldpc_errors"] > 0 else 2 if d["status"]["crc_extracted"] != d["status"]["crc_calculated"] else 4 if d["ok"] else 3)
        if st >= 2:
            assert g_st[k]["crc_extracted"] == d["status"]["crc_extracted"] and g_st[k]["crc_calculated"] == d["status"]["crc_calculated"]
        if st >= 3:
            assert g_st[k]["unpack_status"] == d["status"]["unpack_status"]
        if d["ok"]:
            assert g_msg[k]["text"] == d["msg"]["text"] and g_msg[k]["hash"] == d["msg"]["hash"]
    assert int(nres[s]) == o["n"]
    assert view(res[s], result_dtype).tobytes() == o["results"].tobytes(), "decoder_results[] incl. gaps left by non-CQ messages"
    n = o["n"]
    gu = view(umsg[s], msg_dtype)[:n]
    assert np.array_equal(gu["text"], o["msgs"]["text"]) and np.array_equal(gu["hash"], o["msgs"]["hash"])  # (byte 25 is struct padding)
    assert bits_equal(ufreq[s, :n].cpu().numpy(), o["freq_hz"]) and np.array_equal(uscore[s, :n].cpu().numpy(), o["score"])
    return o


@pytest.fixture(params=[0, 1], ids=["bp_nodes", "bp_edges"])
def bp_variant(request, pkg):
    """Both belief-propagation kernels (ft8b200_set_decode_variant): node-centred (default) and edge-centred."""
    pkg.set_decode_variant(request.param)
    yield request.param
    pkg.set_decode_variant(0)


def test_inlined_pade_division_is_ieee_over_all_floats(ctx):
    """The node-centred kernel inlines the in-range instruction sequence of div.rn.f32 into fast_tanh / fast_atanh
    (ldpc.c:220-251).  Device sweep over all 2^32 float bit patterns against the same expressions with the full IEEE
    division: not one result may differ (the full division itself is what the bit-exact edge-centred kernel uses)."""
    tanh_bad, atanh_bad, atanh_far_bad, tanh_slow, atanh_slow = ctx.selfcheck_pade()
    assert (tanh_bad, atanh_bad, atanh_far_bad) == (0, 0, 0)
    assert 0 < tanh_slow < 2 ** 32 // 4 and 0 < atanh_slow < 2 ** 32  # the fallback exists and is not the common path


def test_decode_degenerate_llrs(ctx, oracle, slots, bp_variant):
    """Candidates placed by hand where the statistics degenerate: a flat waterfall (all LLRs 0 -> variance 0 -> scale inf ->
    NaN LLRs, decode.c:295-314), a waterfall flat except one row, and candidates hanging over both time edges."""
    flat = np.full(94208, 77, np.uint8)
    one_row = flat.copy().reshape(92, 2, 2, 256)
    one_row[40] = np.random.default_rng(4).integers(0, 256, size=(2, 2, 256), dtype=np.uint8)
    real = oracle.waterfall(*slots[0])
    mags = np.stack([flat, one_row.reshape(-1), real])
    cands = np.zeros((3, ctx.K), cand_dtype)
    picks = [(20, 0, 10, 0, 0), (20, -12, 100, 1, 1), (20, 23, 248, 1, 0), (20, 5, 0, 0, 1), (20, -7, 33, 0, 0), (20, 17, 200, 1, 1)]
    for s in range(3):
        for q, pk in enumerate(picks):
            cands[s, q] = pk
    ncand = np.full(3, len(picks), np.int32)
    d_mag = torch.from_numpy(mags).to(dev())
    d_cand = torch.from_numpy(cands.view(np.uint8).reshape(3, ctx.K, 8)).to(dev())
    ok, stage, status, msg, plain, llr = ctx.decode(d_mag, d_cand, torch.from_numpy(ncand).to(dev()), want_plain=True, want_llr=True)
    torch.cuda.synchronize()
    g_st = [view(status[s], status_dtype) for s in range(3)]
    saw_nan = False
    for s in range(3):
        for q in range(len(picks)):
            d = oracle.decode(mags[s], cands[s, q])
            gl = llr[s, q].cpu().numpy()
            saw_nan |= bool(np.isnan(d["llr"]).any())
            assert np.array_equal(np.isnan(gl), np.isnan(d["llr"])) and bits_equal(np.nan_to_num(gl), np.nan_to_num(d["llr"])), (s, q)
            assert np.array_equal(plain[s, q].cpu().numpy(), d["plain"]), (s, q)
            assert g_st[s][q]["ldpc_errors"] == d["status"]["ldpc_errors"] and int(ok[s, q]) == d["ok"], (s, q)
    assert saw_nan


@pytest.mark.parametrize("which", ["k120", "k500"])
def test_sync_decode_spots_parity(ctx, ctx500, oracle, slots, which, bp_variant):
    c = ctx if which == "k120" else ctx500
    mags = np.stack([oracle.waterfall(*s) for s in slots])
    rng = np.random.default_rng(3)
    noise = rng.integers(0, 256, size=(2, mags.shape[1]), dtype=np.uint8)           # random bytes: thousands of survivors, heap evictions
    noise[1] = rng.integers(60, 90, size=mags.shape[1], dtype=np.uint8)
    flat = np.full((1, mags.shape[1]), 77, np.uint8)                                 # constant: every score is 0
    mags = np.concatenate([mags, noise, flat])
    d_mag = torch.from_numpy(mags).to(dev())
    cand, ncand = c.find_sync(d_mag)
    ok, stage, status, msg, plain, llr = c.decode(d_mag, cand, ncand, want_plain=True, want_llr=True)
    res, nres, umsg, ufreq, uscore = c.spots(cand, ncand, ok, msg)
    torch.cuda.synchronize()
    decoded = 0
    for s in range(mags.shape[0]):
        o = check_slot_against_oracle(oracle, c, mags[s], cand, ncand, ok, stage, status, msg, plain, llr, res, nres, umsg, ufreq, uscore, s, c.K, c.M)
        decoded += o["n"]
    assert decoded >= 30
    assert int(ncand[-1]) == 0 and int(ncand[len(slots)]) == c.K


def test_min_score_zero_and_small_k(pkg, oracle, slots):
    """min_score = 0 admits every position (35 856 survivors -> the heap replay is exercised to the full)."""
    c = pkg.Context(0, max_candidates=7, max_messages=50, min_score=0)
    mag = oracle.waterfall(*slots[0])
    cand, ncand = c.find_sync(torch.from_numpy(mag[None]).to(dev()))
    o = oracle.find_sync(mag, max_cand=7, min_score=0)
    assert np.array_equal(view(cand[0], cand_dtype)[: int(ncand[0])], o)
    c.close()


@pytest.mark.parametrize("K", [1, 7, 120, 500])
def test_every_position_survives(pkg, oracle, K):
    """min_score below every possible score: all 35 856 positions reach the heap stage (the reference's loop then pushes K
    and, for each later survivor, tests `score > heap[0].score`).  Random bytes, a narrow band of values (ties everywhere) and
    a ramp whose scores grow along the loop order (evictions all the way): candidates identical in content and order."""
    c = pkg.Context(0, max_candidates=K, max_messages=50, min_score=-1000)
    rng = np.random.default_rng(100 + K)
    mags = np.stack([rng.integers(0, 256, 94208, dtype=np.uint8), rng.integers(100, 104, 94208, dtype=np.uint8),
                     np.clip(np.arange(94208) // 400 + rng.integers(0, 24, 94208), 0, 255).astype(np.uint8)])
    cand, ncand = c.find_sync(torch.from_numpy(mags).to(dev()))
    torch.cuda.synchronize()
    for s in range(mags.shape[0]):
        o = oracle.find_sync(mags[s], max_cand=K, min_score=-1000)
        assert int(ncand[s]) == o.size == K
        assert np.array_equal(view(cand[s], cand_dtype)[:K], o), (K, s)
    c.close()


@pytest.mark.parametrize("geometry", ["ft8_3200", "ft8_12k", "ft4", "odd", "narrow"])
def test_selection_paths(pkg, oracle, slots, geometry):
    """The candidate selection on both of its paths.  The score kernels hand every position with
    score >= min_score to the selection as an unordered list of at most 1024 words per slot (sorted back into the reference's
    loop order); a slot with more survivors is scanned in position order instead.  Thresholds are chosen per waterfall so that
    survivor counts fall on either side of 32 (sorted in registers by one warp), of K and of 1024 (and all positions at once);
    slots of both kinds share a launch."""
    rng = np.random.default_rng(77)
    if geometry == "ft8_3200":
        dims, proto, cells = dict(num_blocks=92, num_bins=256, time_osr=2, freq_osr=2), 1, 94208
    elif geometry == "ft8_12k":
        dims, proto, cells = dict(num_blocks=93, num_bins=960, time_osr=2, freq_osr=2), 1, 93 * 4 * 960
    elif geometry == "ft4":
        dims, proto, cells = dict(num_blocks=156, num_bins=288, time_osr=2, freq_osr=2), 0, 156 * 4 * 288
    elif geometry == "odd":      # 7668 positions: not a multiple of 8 (scalar loads in the scan path), freq_osr 3, unaligned rows
        dims, proto, cells = dict(num_blocks=60, num_bins=78, time_osr=1, freq_osr=3), 1, 60 * 3 * 78
    else:                        # one frequency offset per sub-plane, no oversampling
        dims, proto, cells = dict(num_blocks=95, num_bins=8, time_osr=1, freq_osr=1), 1, 95 * 8
    npos = dims["time_osr"] * dims["freq_osr"] * 36 * (dims["num_bins"] - 7)
    mags = [rng.integers(0, 256, cells, dtype=np.uint8), rng.integers(100, 104, cells, dtype=np.uint8), np.zeros(cells, np.uint8),
            np.clip(np.arange(cells) // max(cells // 230, 1) + rng.integers(0, 24, cells), 0, 255).astype(np.uint8)]
    if geometry == "ft8_3200":
        mags.append(oracle.waterfall(*slots[0]))
    mags = np.stack(mags)

    every = [oracle.find_sync(m, max_cand=npos, min_score=-1000, protocol=proto, **dims)["score"] for m in mags]   # all scores of a slot

    def survivors(s, t):
        return int((every[s] >= t).sum())

    # thresholds around the list capacity on the noise waterfall, plus fixed ones
    counts = {t: survivors(0, t) for t in range(1, 200)}
    below = max((t for t in counts if counts[t] <= 1024), key=lambda t: counts[t])
    above = min((t for t in counts if counts[t] > 1024), key=lambda t: counts[t], default=None)
    few = max((t for t in counts if 0 < counts[t] <= 32), key=lambda t: counts[t], default=None)
    assert npos <= 1024 or counts[below] <= 1024 < counts[above]
    thresholds = [t for t in (few, below, above, 10, 1, 0, -1000) if t is not None]
    d_mag = torch.from_numpy(mags).to(dev())
    for K in (1, 50, 120, 1500):
        for t in thresholds:
            c = pkg.Context(0, max_candidates=K, max_messages=50, min_score=t)
            c.set_protocol(proto)
            cand, ncand = c.find_sync(d_mag, **dims)
            torch.cuda.synchronize()
            g = view(cand, cand_dtype)
            for s in range(mags.shape[0]):
                o = oracle.find_sync(mags[s], max_cand=K, min_score=t, protocol=proto, **dims)
                assert int(ncand[s]) == o.size, (geometry, K, t, s)
                assert g[s, : o.size].tobytes() == o.tobytes(), (geometry, K, t, s, survivors(s, t))
            c.close()


@pytest.mark.parametrize("name,src", [("slot_single", "slot_single"), ("slot_crowded_k500", "slot_crowded_k500"), ("slot_crowded_k120", "slot_crowded_k500")])
def test_golden_slots_on_gpu(ctx, ctx500, name, src):
    """The committed outputs of the UNMODIFIED reference, stage by stage, without the oracle in between."""
    g, gi = golden(name), golden(src)
    c = ctx500 if int(g["kmax"]) == 500 else ctx
    d_i = torch.from_numpy(gi["i"][None]).to(dev()); d_q = torch.from_numpy(gi["q"][None]).to(dev())
    mag = c.waterfall(d_i, d_q)
    assert np.array_equal(mag[0].cpu().numpy(), g["wf"])
    cand, ncand = c.find_sync(mag)
    assert np.array_equal(view(cand[0], cand_dtype)[: int(ncand[0])], g["cands"].view(cand_dtype))
    ok, stage, status, msg, plain, llr = c.decode(mag, cand, ncand, want_plain=True, want_llr=True)
    n = int(ncand[0])
    assert np.array_equal(ok[0, :n].cpu().numpy(), g["dec_ok"].astype(np.uint8))
    assert np.array_equal(plain[0, :n].cpu().numpy(), g["plain"])
    assert bits_equal(llr[0, :n].cpu().numpy(), g["llr"])
    g_st = view(status[0], status_dtype)[:n]
    assert np.array_equal(g_st["ldpc_errors"], g["dec_status"].view(status_dtype)["ldpc_errors"].reshape(-1))
    okm = g["dec_ok"].astype(bool)
    assert view(msg[0], msg_dtype)[:n][okm].tobytes() == g["dec_msg"].view(msg_dtype).reshape(-1)[okm].tobytes()
    c.process_slots(d_i, d_q)
    res, nres = c.fetch_results(1)
    assert nres[0] == int(g["n"]) and res[0].tobytes() == g["results"].tobytes()


def test_process_slots_batch(ctx, oracle, slots):
    """BASELINE config #4 in miniature: a batch of independent slots through the fused path."""
    reps = 5
    hi = np.stack([s[0] for s in slots] * reps); hq = np.stack([s[1] for s in slots] * reps)
    res, nres = ctx.process_slots_host(hi, hq)
    for s, (i_s, q_s) in enumerate(slots):
        o = oracle.subsystem(i_s, q_s)
        for r in range(reps):
            assert nres[s + r * len(slots)] == o["n"]
            assert res[s + r * len(slots)].tobytes() == o["results"].tobytes()


# ------------------------------------------------------------------------------------------- drop-in entry points
def test_dropin_ft8_subsystem(pkg, oracle, slots):
    """ft8_subsystem() with the reference's signature on host arrays (rtlsdr_ft8d.h:164)."""
    for i_s, q_s in slots[:3]:
        o = oracle.subsystem(i_s, q_s)
        dec = np.zeros(50, result_dtype)
        dec["call"] = b"STALE"
        decodes, n = pkg.ft8_subsystem(i_s, q_s, dec)
        assert n == o["n"]
        for k in range(50):
            if o["results"][k]["call"]:
                assert decodes[k].tobytes() == o["results"][k].tobytes()
            else:  # the reference leaves non-CQ slots untouched (rtlsdr_ft8d.c:1509-1520)
                assert decodes[k]["call"] == b"STALE"


def test_dropin_find_sync_and_decode(pkg, oracle, slots):
    """ft8_find_sync()/ft8_decode() with waterfall_t/candidate_t/message_t/decode_status_t in HOST memory."""
    mag = oracle.waterfall(*slots[1])
    heap = pkg.ft8_find_sync(mag, 120, 10)
    o = oracle.find_sync(mag, 120, 10)
    assert np.array_equal(heap.view(cand_dtype), o)
    n_ok = 0
    for cd in o[:40]:
        ok, msg, st = pkg.ft8_decode(mag, cd, 20)
        d = oracle.decode(mag, cd)
        assert ok == bool(d["ok"])
        assert st.tobytes() == d["status"].tobytes(), "unwritten status fields must stay unwritten (0xA5 pattern)"
        if ok:
            assert msg.tobytes() == d["msg"].tobytes()
            n_ok += 1
    assert n_ok >= 3
    # a candidate find_sync never returned + a different iteration count: served by the uncached single-candidate path
    cd = np.array([(11, 3, 100, 1, 0)], cand_dtype)[0]
    for iters in (20, 5):
        ok, msg, st = pkg.ft8_decode(mag, cd, iters)
        d = oracle.decode(mag, cd, max_iters=iters)
        assert ok == bool(d["ok"]) and st.tobytes() == d["status"].tobytes()
    ok, msg, st = pkg.ft8_decode(mag, o[0], 3)
    d = oracle.decode(mag, o[0], max_iters=3)
    assert ok == bool(d["ok"]) and st.tobytes() == d["status"].tobytes()


# ------------------------------------------------------------------------------------------- receiver streams
def test_stream_callback_matches_reference_chunking(pkg, ctx, oracle, raw_slot):
    """rtlsdr_callback() fed as librtlsdr does (65536-byte buffers), filter state carried across calls and across
    the 15 s buffer flip (rtlsdr_ft8d.c:80-86, 1339-1354), against the oracle's sample-by-sample streaming decimator."""
    st = pkg.Stream(ctx)
    rng = np.random.default_rng(8)
    second = rng.integers(0, 256, size=65536 * 300 + 8 * 5, dtype=np.uint8)  # second slot: random bytes, ragged end
    o_state = oracle.new_decim()
    oi1, oq1 = [], []
    for o in range(0, raw_slot.size, 65536):
        chunk = raw_slot[o:o + 65536]
        st.callback(chunk)
        a, b = oracle.decim_feed(o_state, chunk, 64)
        oi1.append(a); oq1.append(b)
    oi1 = np.concatenate(oi1); oq1 = np.concatenate(oq1)
    assert st.count() == oi1.size == 47936
    st.flip()
    gi, gq, n = st.fetch()
    assert n == oi1.size
    assert bits_equal(gi[:n], oi1) and bits_equal(gq[:n], oq1) and not gi[n:].any()
    res, nres = st.decode()  # decoder(): >= 12 s of samples, condition, ft8_subsystem
    ic, qc, _ = oracle.condition(gi, gq, n)
    o = oracle.subsystem(ic, qc)
    assert nres == o["n"] >= 1 and res.tobytes() == o["results"].tobytes()
    # second slot continues from the first slot's filter state; odd call sizes (multiples of 8 bytes)
    oi2, oq2 = [], []
    sizes = [8, 16, 1496, 65536 - 1520] + [65536] * 299 + [40]
    o = 0
    for sz in sizes:
        chunk = second[o:o + sz]; o += sz
        st.callback(chunk)
        a, b = oracle.decim_feed(o_state, chunk, 64)
        oi2.append(a); oq2.append(b)
    assert o == second.size
    oi2 = np.concatenate(oi2); oq2 = np.concatenate(oq2)
    assert st.count() == oi2.size
    st.flip()
    gi, gq, n = st.fetch()
    assert n == oi2.size and bits_equal(gi[:n], oi2) and bits_equal(gq[:n], oq2)
    res, nres = st.decode()
    assert nres == -1, "fewer than 12 s of samples: decoder() skips the slot (rtlsdr_ft8d.c:235-238)"
    st.close()


def test_callback_thread_and_decoder_thread(pkg, ctx, raw_slot):
    """The daemon's two threads: librtlsdr's USB thread in rtlsdr_callback() (rtlsdr_ft8d.c:214) while the decoder thread works on
    the slot that was just flipped out (:274, decoder() :221-285).  Slot n is decoded -- repeatedly -- while slot n+1 arrives; the
    spots of both slots and the samples of slot n+1 are those of the same calls made one after the other."""
    import threading
    second = np.roll(raw_slot, 2 * 1000)   # the same signal 1000 samples later: another slot with a message in it

    def feed(st, buf):
        for o in range(0, buf.size, 65536):
            st.callback(buf[o:o + 65536])

    serial = pkg.Stream(ctx)
    feed(serial, raw_slot); serial.flip()
    want_a = serial.decode()
    feed(serial, second); serial.flip()
    want_i, want_q, want_n = serial.fetch()
    want_b = serial.decode()
    serial.close()
    assert want_a[1] >= 1 and want_b[1] >= 1

    st = pkg.Stream(ctx)
    feed(st, raw_slot); st.flip()
    usb = threading.Thread(target=feed, args=(st, second))
    usb.start()
    rounds = 0
    while usb.is_alive() or rounds < 3:   # ctypes releases the GIL in every call: the two threads really overlap
        res, n = st.decode()
        assert n == want_a[1] and res.tobytes() == want_a[0].tobytes()
        rounds += 1
    usb.join()
    st.flip()
    gi, gq, n = st.fetch()
    assert n == want_n and bits_equal(gi, want_i) and bits_equal(gq, want_q)
    res, n = st.decode()
    assert n == want_b[1] and res.tobytes() == want_b[0].tobytes()
    st.close()


def test_default_stream_callback(pkg, oracle):
    """ctx == NULL: the process-wide stream, like the reference's function-static state."""
    rng = np.random.default_rng(9)
    buf = rng.integers(0, 256, size=65536 * 3, dtype=np.uint8)
    for o in range(0, buf.size, 65536):
        pkg.rtlsdr_callback(buf[o:o + 65536])
    # an odd-sized buffer is rejected loudly, not half-processed
    pkg.rtlsdr_callback(buf[:12])


# ------------------------------------------------------------------------------------------- 12 kHz monitor path (ft8_lib decode_ft8.c)
@pytest.fixture(scope="module")
def audio():
    sigs = [(ft8enc.tones(ft8enc.pack_std("CQ", "K1JT", "FN20")), 1200.0, 0.5, 0.1),
            (ft8enc.tones(ft8enc.pack_std("K1ABC", "W9XYZ", "-15")), 2100.0, 1.1, 0.05),
            (ft8enc.tones(ft8enc.pack_std("W9XYZ", "K1ABC", "RR73")), 310.0, 0.1, 0.03)]
    return synth.audio_12k(sigs, 3)


def test_monitor_batch_waterfall(ctx, oracle, audio):
    """3840-point real FFT frames (kiss_fftr arithmetic), Hann window, 93 x 2 x 2 x 960 cells: identical bytes."""
    ref, info, _ = oracle.monitor_waterfall(audio)
    quiet = np.zeros_like(audio)
    short = np.zeros_like(audio); short[:50_000] = audio[:50_000]
    batch = torch.from_numpy(np.stack([audio, quiet, short])).to(dev())
    mag, nb = ctx.monitor_waterfall(batch)
    assert nb == 93 == int(info[4])
    mag = mag.cpu().numpy()
    assert np.array_equal(mag[0], ref), f"{int((mag[0] != ref).sum())} cells differ"
    assert np.array_equal(mag[1], oracle.monitor_waterfall(quiet)[0])
    assert np.array_equal(mag[2], oracle.monitor_waterfall(short)[0])
    # a shorter recording: fewer blocks
    mag2, nb2 = ctx.monitor_waterfall(batch[:1, :1920 * 40 + 100].contiguous())
    ref2, info2, _ = oracle.monitor_waterfall(audio[:1920 * 40 + 100])
    assert nb2 == 40 == int(info2[4]) and np.array_equal(mag2[0].cpu().numpy()[: ref2.size], ref2)


def test_monitor_dropin_and_decode(pkg, oracle, audio):
    """monitor_init/monitor_process/monitor_reset + ft8_find_sync/ft8_decode on the 960-bin waterfall, as decode_ft8 main() uses them."""
    mon = pkg.Monitor(12000, 2, 2, 1)
    assert (mon.me.block_size, mon.me.subblock_size, mon.me.nfft, mon.me.wf.max_blocks, mon.me.wf.num_bins) == (1920, 960, 3840, 93, 960)
    for o in range(0, audio.size - 1920 + 1, 1920):
        mon.process(audio[o:o + 1920])
    mon.process(audio[:1920])  # 94th block: silently ignored
    ref, info, ref_max = oracle.monitor_waterfall(audio)
    assert mon.me.wf.num_blocks == 93
    assert np.array_equal(mon.mag(), ref)
    assert np.float32(mon.me.max_mag) == np.float32(ref_max)
    heap = mon.find_sync(120, 10)
    o_heap = oracle.find_sync(ref, 120, 10, num_blocks=93, num_bins=960)
    assert np.array_equal(heap.view(cand_dtype), o_heap)
    texts = set()
    for cd in o_heap:
        ok, msg, st = mon.decode(cd, 20)
        d = oracle.decode(ref, cd, num_blocks=93, num_bins=960)
        assert ok == bool(d["ok"]) and st.tobytes() == d["status"].tobytes()
        if ok:
            assert msg["text"] == d["msg"]["text"]
            texts.add(msg["text"].decode())
    assert {"CQ K1JT FN20", "K1ABC W9XYZ -15"} <= texts
    mon.reset()
    assert mon.me.wf.num_blocks == 0
    mon.process(audio[:1920])  # history is NOT cleared by reset (decode_ft8.c:220-224)
    assert mon.me.wf.num_blocks == 1
    mon.close()


def test_monitor_deferred_mode(pkg, oracle, audio):
    """ft8b200_monitor_set_deferred: monitor_process() only appends, ONE launch transforms the pending blocks when the waterfall is
    needed (ft8_find_sync / ft8_decode / flush / reset / free).  Same bytes, max_mag, last_frame, candidates and messages as the
    block-by-block mode, also when the two modes are mixed and across a reset."""
    strict, lazy = pkg.Monitor(12000, 2, 2, 1), pkg.Monitor(12000, 2, 2, 1)
    lazy.set_deferred(True)
    blocks = [audio[o:o + 1920] for o in range(0, audio.size - 1920 + 1, 1920)]
    for b in blocks[:40]:
        strict.process(b); lazy.process(b)
    assert lazy.me.wf.num_blocks == strict.me.wf.num_blocks == 40
    assert lazy.flush() == 40 and lazy.flush() == 0
    assert np.array_equal(lazy.mag(), strict.mag()) and np.float32(lazy.me.max_mag) == np.float32(strict.me.max_mag)
    lazy.set_deferred(False)
    for b in blocks[40:50]:
        strict.process(b); lazy.process(b)          # block by block again
    lazy.set_deferred(True)
    for b in blocks[50:]:
        strict.process(b); lazy.process(b)
    lazy.process(blocks[0]); strict.process(blocks[0])   # 94th block: ignored in both modes
    heap_l = lazy.find_sync(120, 10)                # flushes by itself
    heap_s = strict.find_sync(120, 10)
    assert np.array_equal(heap_l, heap_s) and np.array_equal(lazy.mag(), strict.mag()) and lazy.me.wf.num_blocks == 93
    assert np.float32(lazy.me.max_mag) == np.float32(strict.me.max_mag)
    lf_l = np.ctypeslib.as_array(__import__("ctypes").cast(lazy.me.last_frame, __import__("ctypes").POINTER(__import__("ctypes").c_float)), shape=(3840,))
    lf_s = np.ctypeslib.as_array(__import__("ctypes").cast(strict.me.last_frame, __import__("ctypes").POINTER(__import__("ctypes").c_float)), shape=(3840,))
    assert np.array_equal(lf_l, lf_s)
    ref, _, _ = oracle.monitor_waterfall(audio)
    assert np.array_equal(lazy.mag(), ref)
    for cd in heap_s[:10]:
        a, b = lazy.decode(cd, 20), strict.decode(cd, 20)
        assert a[0] == b[0] and a[1].tobytes() == b[1].tobytes() and a[2].tobytes() == b[2].tobytes()
    # reset with blocks pending: they still shape the history; then the next recording
    lazy.reset(); strict.reset()
    for b in blocks[:5]:
        strict.process(b); lazy.process(b)
    lazy.reset(); strict.reset()
    for b in blocks[5:9]:
        strict.process(b); lazy.process(b)
    lazy.flush()
    assert np.array_equal(lazy.mag(), strict.mag()) and lazy.me.wf.num_blocks == 4
    lazy.process(blocks[9])                         # freed with a block pending
    lazy.close(); strict.close()


def test_grouped_overlap_gives_identical_results(ctx, raw_slot):
    """Large raw batches are split into slot groups whose back end overlaps the next group's decimator on a side
    stream; that must not change a single byte, and every slot must equal its stand-alone result."""
    B = 20
    big = torch.from_numpy(raw_slot).to(dev()).repeat(B, 1).contiguous()
    big[3] = 0x80                      # a silent slot
    big[7, : 2 * 751 * 9000] = 0x80    # signal starts late: different peak, same message
    big[11] = big[11].flip(0)          # garbage
    outs = []
    for ov in (0, 2, 5):
        ctx.set_overlap(ov)
        ctx.process_raw(big, B)
        outs.append(ctx.fetch_results(B))
    ctx.set_overlap(0)
    for res, n in outs[1:]:
        assert np.array_equal(n, outs[0][1]) and res.tobytes() == outs[0][0].tobytes()
    res, n = outs[0]
    for s in (0, 3, 7, 11, 19):
        ctx.process_raw(big[s:s + 1], 1)
        r1, n1 = ctx.fetch_results(1)
        assert n1[0] == n[s] and r1[0].tobytes() == res[s].tobytes()
    assert n[0] >= 1 and n[3] == 0


# ------------------------------------------------------------------------------------------- pipelined executor
def _mixed_batch(raw_slot, B):
    big = torch.from_numpy(raw_slot).to(dev()).repeat(B, 1).contiguous()
    big[1] = 0x80                      # a silent slot
    big[2, : 2 * 751 * 9000] = 0x80    # signal starts late: different peak, same message
    big[3] = big[3].flip(0)            # garbage
    return big


@pytest.mark.parametrize("depth,serial,variant", [(1, False, 0), (2, False, 0), (2, True, 0), (3, False, 5), (2, True, 1)])
def test_pipe_matches_unpipelined(pkg, ctx, oracle, raw_slot, depth, serial, variant):
    """ft8b200_pipe_t (several batches in flight, back end on a side stream, optional bulk-copy decimator) returns exactly
    the records of ft8b200_process_raw + ft8b200_fetch_results, batch after batch, for device and for host input; the
    first slot is also checked against the CPU oracle."""
    B = 6
    big = _mixed_batch(raw_slot, B)
    torch.cuda.synchronize()
    ctx.process_raw(big, B)
    ref_res, ref_n = ctx.fetch_results(B)
    oi, oq = oracle.decimate_slot(raw_slot)
    ri = np.zeros(48000, np.float32); rq = np.zeros(48000, np.float32)
    ri[:oi.size] = oi; rq[:oq.size] = oq
    o = oracle.subsystem(*oracle.condition(ri, rq, oi.size)[:2])
    assert ref_n[0] == o["n"] >= 1 and ref_res[0].tobytes() == o["results"].tobytes()

    pipe = pkg.Pipe(0, depth)
    pipe.set_mode(serial, variant)
    outs = []
    sizes = [B, 2, B, 1, 4, B, 3]
    for n in sizes:
        if pipe.in_flight() == pipe.depth:
            outs.append(pipe.collect(B))
        pipe.submit(big[:n], n)
    with pytest.raises(pkg.Ft8Error):   # FT8B200_EBUSY only when every lane is in flight
        while True:
            pipe.submit(big[:1], 1)
            sizes.append(1)
    while pipe.in_flight():
        outs.append(pipe.collect(B))
    assert [len(o[1]) for o in outs] == sizes
    for (res, n), k in zip(outs, sizes):
        assert np.array_equal(n, ref_n[:k]) and res.tobytes() == ref_res[:k].tobytes()
    with pytest.raises(pkg.Ft8Error):
        pipe.collect(B)                 # nothing in flight

    host = big[:4].cpu().numpy()
    houts = []
    for n in (4, 1, 3, 4):
        if pipe.in_flight() == pipe.depth:
            houts.append(pipe.collect(B))
        pipe.submit_host(host[:n], n)
    while pipe.in_flight():
        houts.append(pipe.collect(B))
    for (res, n), k in zip(houts, (4, 1, 3, 4)):
        assert np.array_equal(n, ref_n[:k]) and res.tobytes() == ref_res[:k].tobytes()
    assert pipe.launches() > 0
    pipe.close()


@pytest.mark.parametrize("depth,back_sms", [(2, 8), (3, 40), (2, 100)])
def test_pipe_sm_partition(pkg, ctx, oracle, raw_slot, depth, back_sms):
    """ft8b200_pipe_set_partition: decimator of batch n+1 and back end of batch n on disjoint SM sets (green contexts).
    Same records as the unpipelined call for device and host input, through partition changes, removal and mode switches;
    the first slot is also checked against the CPU oracle."""
    B = 6
    big = _mixed_batch(raw_slot, B)
    torch.cuda.synchronize()
    ctx.process_raw(big, B)
    ref_res, ref_n = ctx.fetch_results(B)
    oi, oq = oracle.decimate_slot(raw_slot)
    ri = np.zeros(48000, np.float32); rq = np.zeros(48000, np.float32)
    ri[:oi.size] = oi; rq[:oq.size] = oq
    o = oracle.subsystem(*oracle.condition(ri, rq, oi.size)[:2])
    assert ref_n[0] == o["n"] >= 1 and ref_res[0].tobytes() == o["results"].tobytes()

    pipe = pkg.Pipe(0, depth)
    host = big[:4].cpu().numpy()

    def rounds():
        outs, sizes = [], [B, 2, B, 1, 4, B, 3]
        for n in sizes:
            if pipe.in_flight() == pipe.depth:
                outs.append(pipe.collect(B))
            pipe.submit(big[:n], n)
        for n in (4, 1, 3):
            if pipe.in_flight() == pipe.depth:
                outs.append(pipe.collect(B))
            pipe.submit_host(host[:n], n)
            sizes.append(n)
        while pipe.in_flight():
            outs.append(pipe.collect(B))
        assert [len(o[1]) for o in outs] == sizes
        for (res, n), k in zip(outs, sizes):
            assert np.array_equal(n, ref_n[:k]) and res.tobytes() == ref_res[:k].tobytes()

    f, b = pipe.set_partition(back_sms)
    assert b >= back_sms and f >= 1 and f + b <= torch.cuda.get_device_properties(0).multi_processor_count
    rounds()
    pipe.submit(big[:1], 1)
    with pytest.raises(pkg.Ft8Error):   # not while a batch is in flight
        pipe.set_partition(16)
    pipe.collect(B)
    f2, b2 = pipe.set_partition(16)     # re-partition
    assert b2 >= 16
    rounds()
    assert pipe.set_partition(0) == (0, 0)   # back to time sharing
    rounds()
    pipe.set_partition(back_sms)
    pipe.set_mode(True)                 # a mode switch drops the partition
    rounds()
    with pytest.raises(pkg.Ft8Error):
        pipe.set_partition(10 ** 6)
    pipe.close()
    one = pkg.Pipe(0, 1)
    with pytest.raises(pkg.Ft8Error):   # nothing to overlap with a single lane
        one.set_partition(16)
    one.close()


def test_pipe_sm_partition_layouts(pkg, ctx, raw_slot):
    """ft8b200_pipe_set_partition(back_sms + 1000 * layout): the back partition composed of the driver's 8-SM groups in every
    layout is disjoint from the front partition, both cover the GPU, and the records are those of the unpipelined call."""
    B = 4
    big = _mixed_batch(raw_slot, B)
    torch.cuda.synchronize()
    ctx.process_raw(big, B)
    ref_res, ref_n = ctx.fetch_results(B)
    n_sm = torch.cuda.get_device_properties(0).multi_processor_count
    pipe = pkg.Pipe(0, 3)
    whole = set(pipe.partition_smids(0))
    assert len(whole) == n_sm
    seen = set()
    for layout in range(0, 7):
        f, b = pipe.set_partition(32 + 1000 * layout)
        assert b == 32 and f == n_sm - 32
        back, front = set(pipe.partition_smids(1)), set(pipe.partition_smids(0))
        assert len(back) == 32 and len(front) == f and not (back & front) and (back | front) == whole
        seen.add(tuple(sorted(back)))
        for _ in range(4):
            if pipe.in_flight() == pipe.depth:
                res, n = pipe.collect(B)
                assert np.array_equal(n, ref_n) and res.tobytes() == ref_res.tobytes()
            pipe.submit(big, B)
        while pipe.in_flight():
            res, n = pipe.collect(B)
            assert np.array_equal(n, ref_n) and res.tobytes() == ref_res.tobytes()
    assert len(seen) >= 3   # the layouts really are different SM sets
    # back ends of consecutive batches chained (the autotune's probe mode): same records
    pipe.set_partition(24)
    pipe.set_back_chain(True)
    for _ in range(7):
        if pipe.in_flight() == pipe.depth:
            res, n = pipe.collect(B)
            assert np.array_equal(n, ref_n) and res.tobytes() == ref_res.tobytes()
        pipe.submit(big, B)
    while pipe.in_flight():
        res, n = pipe.collect(B)
        assert np.array_equal(n, ref_n) and res.tobytes() == ref_res.tobytes()
    pipe.submit(big, B)
    with pytest.raises(pkg.Ft8Error):   # not while a batch is in flight
        pipe.set_back_chain(False)
    pipe.collect(B)
    pipe.set_back_chain(False)
    with pytest.raises(pkg.Ft8Error):
        pipe.set_partition(7032)
    with pytest.raises(pkg.Ft8Error):
        pipe.set_partition(2000)
    pipe.close()


@pytest.mark.parametrize("variant", [1, 2, 3, 4, 5, 6])
def test_bulk_copy_decimator_variants(pkg, oracle, variant):
    """The persistent cp.async.bulk + mbarrier cic_block_sums kernel (every ring shape) is bit-identical to the oracle,
    including batches smaller than one CTA's consumer count and ragged tails handled by the generic kernel."""
    c = pkg.Context(0)
    c.set_decimator_variant(variant)
    rng = np.random.default_rng(40 + variant)
    for nstreams, nbytes in ((1, 12016 * 3), (3, 12016 * 40 + 1502 * 4), (2, 12016 * 700 + 8 * 11), (5, 12016 * 17)):
        stride = (nbytes + 15) // 16 * 16
        iq = rng.integers(0, 256, size=(nstreams, stride), dtype=np.uint8)
        d_i, d_q, cnt, peak, y2 = c.decimate(torch.from_numpy(iq).to(dev()), nstreams, nbytes, stride, want_y2=True)
        for s in range(nstreams):
            oi, oq, oy2i, oy2q = oracle_decim(oracle, iq[s, :nbytes])
            n = int(cnt[s])
            assert n == oi.size
            assert bits_equal(d_i[s, :n].cpu().numpy(), oi) and bits_equal(d_q[s, :n].cpu().numpy(), oq)
            gy = y2[s].cpu().numpy()
            assert np.array_equal(gy[:n, 0], oy2i) and np.array_equal(gy[:n, 1], oy2q)
    c.close()


# ------------------------------------------------------------------------------------------- config #5: continuous streams
def oracle_stream_slots(oracle, iq, slots, bytes_per_slot, chunk):
    """The reference's behaviour for one receiver: rtlsdr_callback() in `chunk`-byte calls with persistent filter state,
    rx buffer flipped every bytes_per_slot bytes (between two callbacks, like main()'s 15 s timer)."""
    st = oracle.new_decim()
    out = []
    for g in range(slots):
        parts = [oracle.decim_feed(st, iq[o:o + chunk], chunk // 2 // 751 + 2, want_y2=True)
                 for o in range(g * bytes_per_slot, (g + 1) * bytes_per_slot, chunk)]
        out.append(tuple(np.concatenate([p[k] for p in parts]) for k in range(4)))
    return out


@pytest.mark.parametrize("bytes_per_slot,slots,chunk", [(1502 * 400 + 600, 5, 6200), (12016 * 50, 3, 12016 * 10), (1502 * 92 + 8, 8, 23032)])
def test_stream_batches_cross_slot_boundaries(ctx, oracle, bytes_per_slot, slots, chunk):
    """BASELINE config #5 at reduced size: consecutive slots of continuous streams, decimated in ONE launch, equal the
    reference's callback-by-callback outputs with the filter state carried across the slot flip: per-slot sample counts
    (the decimation phase drifts), integer CIC outputs, float samples, peaks."""
    assert bytes_per_slot % chunk == 0
    n_streams = 3
    rng = np.random.default_rng(bytes_per_slot)
    stride = (slots * bytes_per_slot + 15) // 16 * 16 + 32
    iq = rng.integers(0, 256, size=(n_streams, stride), dtype=np.uint8)
    iq[1] = rng.integers(120, 137, size=stride, dtype=np.uint8)
    d_i, d_q, cnt, peak, y2 = ctx.decimate_streams(torch.from_numpy(iq).to(dev()), n_streams, slots, bytes_per_slot,
                                                   bytes_per_stream=slots * bytes_per_slot, stride=stride, want_y2=True)
    counts = set()
    for s in range(n_streams):
        ref = oracle_stream_slots(oracle, iq[s], slots, bytes_per_slot, chunk)
        for g, (oi, oq, oy2i, oy2q) in enumerate(ref):
            r = s * slots + g
            n = int(cnt[r])
            counts.add(n)
            assert n == oi.size
            gy = y2[r].cpu().numpy()
            assert np.array_equal(gy[:n, 0], oy2i) and np.array_equal(gy[:n, 1], oy2q)
            assert bits_equal(d_i[r, :n].cpu().numpy(), oi) and bits_equal(d_q[r, :n].cpu().numpy(), oq)
            assert not d_i[r, n:].any() and not d_q[r, n:].any()
            assert float(peak[r]) == float(max(np.abs(oi).max(initial=0), np.abs(oq).max(initial=0)))
    if slots * (bytes_per_slot // 2 % 751) >= 751:
        assert len(counts) > 1, "the decimation phase drifts by more than one output over these slots, so the per-slot counts must differ"


def test_streams_full_size_whole_path(ctx, oracle, raw_slot):
    """Full-size slots: 2 receivers x 3 consecutive 72 MB slots in one call.  Slot 0 of each stream equals the fresh-state
    single-slot result; later slots equal the oracle's continuing stream (47 936 / 47 937 samples) end to end."""
    slots, n_streams = 3, 2
    one = torch.from_numpy(raw_slot).to(dev())
    big = one.repeat(n_streams, slots).contiguous()
    big[1, 72_000_000:2 * 72_000_000] = 0x80      # a silent middle slot on the second receiver
    torch.cuda.synchronize()
    d_i, d_q, cnt, peak, _ = ctx.decimate_streams(big, n_streams, slots, 72_000_000)
    first = [0 if g == 0 else -((750 - g * 36_000_000) // 751) for g in range(slots + 1)]   # ceil((g*N - 750) / 751)
    first[slots] = min(first[slots], slots * 36_000_000 // 751)
    assert cnt.cpu().tolist() == [first[g + 1] - first[g] for g in range(slots)] * n_streams  # 47 936 each (47 937 first appears in slot 11)
    ctx.process_raw_streams(big, n_streams, slots)
    res, n = ctx.fetch_results(n_streams * slots)
    ctx.process_raw(one[None], 1)
    r1, n1 = ctx.fetch_results(1)
    assert n[0] == n[3] == n1[0] >= 1 and res[0].tobytes() == res[3].tobytes() == r1[0].tobytes()
    assert n[4] == 0
    # oracle: continue the first receiver's stream through its second slot
    st = oracle.new_decim()
    host = big[0].cpu().numpy()
    for g in range(2):
        parts = [oracle.decim_feed(st, host[o:o + 600_000], 600_000 // 2 // 751 + 2) for o in range(g * 72_000_000, (g + 1) * 72_000_000, 600_000)]
    oi = np.concatenate([p[0] for p in parts]); oq = np.concatenate([p[1] for p in parts])
    assert oi.size == int(cnt[1])
    assert bits_equal(d_i[1, :oi.size].cpu().numpy(), oi) and bits_equal(d_q[1, :oq.size].cpu().numpy(), oq)
    ri = np.zeros(48000, np.float32); rq = np.zeros(48000, np.float32)
    ri[:oi.size] = oi; rq[:oq.size] = oq
    o = oracle.subsystem(*oracle.condition(ri, rq, oi.size)[:2])
    assert n[1] == o["n"] and res[1].tobytes() == o["results"].tobytes()


@pytest.mark.parametrize("partition", [0, 32])
def test_pipe_streams_match_unpipelined(pkg, ctx, raw_slot, partition):
    """ft8b200_pipe_submit_streams (BASELINE config #5 through the pipelined executor): batches of receiver streams of consecutive
    slots, several in flight, with and without the SM partition -> the records of ft8b200_process_raw_streams, in (stream, slot)
    order, whatever the batch size."""
    slots, n_streams = 2, 3
    one = torch.from_numpy(raw_slot).to(dev())
    big = one.repeat(n_streams, slots).contiguous()
    big[1, 72_000_000:] = 0x80                      # a silent second slot on the second receiver
    big[2, :72_000_000] = big[2, :72_000_000].flip(0)   # another first slot on the third (I/Q swapped and reversed: nothing decodes)
    torch.cuda.synchronize()
    ctx.process_raw_streams(big, n_streams, slots)
    ref_res, ref_n = ctx.fetch_results(n_streams * slots)
    assert ref_n[0] >= 1 and ref_n[3] == 0
    pipe = pkg.Pipe(0, 2)
    if partition:
        pipe.set_partition(partition)
    outs, sizes = [], [3, 1, 2, 3, 1]
    for n in sizes:
        if pipe.in_flight() == pipe.depth:
            outs.append(pipe.collect(n_streams * slots))
        pipe.submit_streams(big[:n], n, slots)
    while pipe.in_flight():
        outs.append(pipe.collect(n_streams * slots))
    assert [len(o[1]) for o in outs] == [n * slots for n in sizes]
    for (res, n), k in zip(outs, sizes):
        assert np.array_equal(n, ref_n[: k * slots]) and res.tobytes() == ref_res[: k * slots].tobytes()
    pipe.close()


# ------------------------------------------------------------------------------------------- FT4
def ft4_audio(seed, n=6):
    rng = np.random.default_rng(seed)
    sigs, texts = [], []
    for k in range(n):
        to, de, ex = synth.random_message(rng)
        if k % 2 == 0:
            to = "CQ"
        sigs.append((ft8enc.tones_ft4(ft8enc.pack_std(to, de, ex)), float(rng.uniform(300.0, 2600.0)), float(0.3 + rng.uniform(0.0, 0.9)),
                     float(rng.uniform(0.03, 0.15))))
        texts.append(f"{to} {de} {ex}")
    return synth.audio_12k(sigs, seed, n_samples=90_000, symbol_period=0.048), texts


def test_ft4_batched(pkg, oracle):
    """FT4 through the batched device API: 7.5 s monitor waterfall (576-sample symbols, 1152-point real FFT, 288 bins),
    ft4_sync_score + top-K, 87x2 LLRs, LDPC, CRC, descrambling, unpack -- all identical to the CPU oracle
    (itself pinned to the unmodified reference by tests/test_oracle_vs_ref.py::test_ft4_vs_reference)."""
    c = pkg.Context(0)
    c.set_protocol(0)
    audios, texts = zip(*[ft4_audio(s) for s in (11, 12, 13)])
    noise = np.random.default_rng(1).standard_normal(90_000).astype(np.float32) * 0.05
    batch = np.stack(list(audios) + [noise])
    mag, nb = c.monitor_waterfall(torch.from_numpy(batch).to(dev()), protocol=0)
    refs = [oracle.monitor_waterfall(a, protocol=0) for a in batch]
    assert nb == int(refs[0][1][4]) == 156
    dims = dict(num_blocks=nb, num_bins=288, time_osr=2, freq_osr=2)
    hm = mag.cpu().numpy()
    for s in range(batch.shape[0]):
        assert np.array_equal(hm[s][: refs[s][0].size], refs[s][0]), f"FT4 waterfall {s}: {(hm[s][:refs[s][0].size] != refs[s][0]).sum()} cells differ"
    cand, ncand = c.find_sync(mag, **dims)
    ok, stage, status, msg, plain, llr = c.decode(mag, cand, ncand, want_plain=True, want_llr=True, **dims)
    torch.cuda.synchronize()
    g_cand, g_ok = view(cand, cand_dtype), ok.cpu().numpy()
    g_st, g_msg = view(status, status_dtype), view(msg, msg_dtype)
    decoded = set()
    for s in range(batch.shape[0]):
        o_heap = oracle.find_sync(refs[s][0], c.K, 10, protocol=0, **dims)
        n = int(ncand[s])
        assert n == o_heap.size and g_cand[s, :n].tobytes() == o_heap.tobytes()
        for k in range(n):
            d = oracle.decode(refs[s][0], o_heap[k], 20, protocol=0, **dims)
            assert bool(g_ok[s, k]) == bool(d["ok"])
            assert np.array_equal(llr[s, k].cpu().numpy().view(np.uint32), d["llr"].view(np.uint32))
            assert np.array_equal(plain[s, k].cpu().numpy(), d["plain"])
            assert g_st[s, k]["ldpc_errors"] == d["status"]["ldpc_errors"]
            if d["ok"]:
                assert g_st[s, k].tobytes() == d["status"].tobytes()
                assert g_msg[s, k]["text"] == d["msg"]["text"] and g_msg[s, k]["hash"] == d["msg"]["hash"]
                decoded.add(d["msg"]["text"].decode())
    assert len(decoded & {t for tt in texts for t in tt}) >= 10
    c.close()


def test_ft4_dropin(pkg, oracle):
    """monitor_init(protocol = PROTO_FT4) / monitor_process / ft8_find_sync / ft8_decode with host structs, as
    `decode_ft8 -ft4` drives them (ft8_lib/decode_ft8.c:240-243,286-330)."""
    audio, texts = ft4_audio(21)
    mon = pkg.Monitor(12000, 2, 2, 0)
    assert (mon.me.block_size, mon.me.subblock_size, mon.me.nfft, mon.me.wf.max_blocks, mon.me.wf.num_bins) == (576, 288, 1152, 156, 288)
    for o in range(0, audio.size - 576 + 1, 576):
        mon.process(audio[o:o + 576])
    ref, info, ref_max = oracle.monitor_waterfall(audio, protocol=0)
    assert mon.me.wf.num_blocks == int(info[4]) and np.array_equal(mon.mag(), ref)
    assert np.float32(mon.me.max_mag) == np.float32(ref_max)
    dims = dict(num_blocks=int(info[4]), num_bins=288, time_osr=2, freq_osr=2, protocol=0)
    heap = mon.find_sync(120, 10)
    o_heap = oracle.find_sync(ref, 120, 10, **dims)
    assert np.array_equal(heap.view(cand_dtype), o_heap)
    got = set()
    for cd in o_heap:
        ok, msg, st = mon.decode(cd, 20)
        d = oracle.decode(ref, cd, 20, **dims)
        assert ok == bool(d["ok"]) and st.tobytes() == d["status"].tobytes()
        if ok:
            assert msg["text"] == d["msg"]["text"]
            got.add(msg["text"].decode())
    assert len(got & set(texts)) >= 4
    mon.close()


# ------------------------------------------------------------------------------------------- recordings on disk
def _write_wav(path, pcm, rate=12000):
    import wave
    with wave.open(str(path), "wb") as w:
        w.setnchannels(1); w.setsampwidth(2); w.setframerate(rate)
        w.writeframes(np.ascontiguousarray(pcm, np.int16).tobytes())


def test_real_recordings_match_reference_stdout(pkg, ctx, tmp_path):
    """Three of the reference's real-world 12 kHz recordings, as WAV files, through ft8b200_decode_wav_files in ONE batch:
    every printed line (score, time, frequency, text, order) equals the stdout of the reference's own `decode_ft8` main()
    stored in tests/golden/recordings_12k.npz."""
    g = golden("recordings_12k")
    paths = []
    for k, pcm in enumerate(g["pcm"]):
        paths.append(str(tmp_path / f"rec{k}.wav"))
        _write_wav(paths[-1], pcm)
    paths.append(str(tmp_path / "missing.wav"))
    outs, status = pkg.decode_wav_files(ctx, paths)
    assert list(status) == [0, 0, 0, -3] and len(outs[3]) == 0
    total = 0
    for k in range(3):
        assert [pkg.format_decoded(r) for r in outs[k]] == str(g["lines"][k]).split("\n"), str(g["names"][k])
        total += len(outs[k])
    assert total == 47


def test_wav_batch_ft4_and_short_files(pkg, ctx, oracle, tmp_path):
    """Synthetic recordings as s16 WAV files: FT4 (7.5 s) and FT8 files of different lengths in one call each,
    against the oracle's decode_ft8 chain on the same quantised samples."""
    a4, _ = ft4_audio(31)
    a4b, _ = ft4_audio(32)
    pcm4 = [np.clip(np.round(a * 20000), -32768, 32767).astype(np.int16) for a in (a4, a4b)]
    paths = []
    for k, p in enumerate(pcm4):
        paths.append(tmp_path / f"ft4_{k}.wav"); _write_wav(paths[-1], p)
    outs, status = pkg.decode_wav_files(ctx, [str(p) for p in paths], protocol=0)
    n = 0
    for k, p in enumerate(pcm4):
        want = oracle.decode_ft8_lines(p.astype(np.float32) / np.float32(32768.0), 12000, protocol=0)
        assert [pkg.format_decoded(r) for r in outs[k]] == want
        n += len(want)
    assert n >= 6
    sigs = [(ft8enc.tones(ft8enc.pack_std("CQ", "K1JT", "FN20")), 1200.0, 0.5, 0.1), (ft8enc.tones(ft8enc.pack_std("K1ABC", "W9XYZ", "-15")), 2100.0, 1.1, 0.05)]
    a8 = synth.audio_12k(sigs, 3)
    pcm8 = np.clip(np.round(a8 * 20000), -32768, 32767).astype(np.int16)
    lens = [180_000, 180_000, 150_000, 1000]
    paths = []
    for k, L in enumerate(lens):
        paths.append(tmp_path / f"ft8_{k}.wav"); _write_wav(paths[-1], pcm8[:L])
    outs, status = pkg.decode_wav_files(ctx, [str(p) for p in paths])
    for k, L in enumerate(lens):
        want = oracle.decode_ft8_lines(pcm8[:L].astype(np.float32) / np.float32(32768.0), 12000)
        assert [pkg.format_decoded(r) for r in outs[k]] == want
    assert len(outs[0]) >= 2 and len(outs[3]) == 0
    # files of different sample rates in one call: each is taken at its own rate, like decode_ft8 does (decode_ft8.c:277-285);
    # the 16 kHz file is the same audio played a third faster (so a signal 1/3 higher and shorter: whatever decode_ft8 makes of it)
    rates = [12000, 16000, 16000, 12000, 11025]
    paths = []
    for k, r in enumerate(rates):
        paths.append(tmp_path / f"mixed_{k}.wav"); _write_wav(paths[-1], pcm8[:150_000 + 1000 * k], r)
    outs, status = pkg.decode_wav_files(ctx, [str(p) for p in paths])
    for k, r in enumerate(rates):
        want = oracle.decode_ft8_lines(pcm8[:150_000 + 1000 * k].astype(np.float32) / np.float32(32768.0), r)
        assert [pkg.format_decoded(x) for x in outs[k]] == want, (k, r)
    assert len(outs[0]) >= 2


def test_iq_and_c2_files(pkg, ctx, oracle, slots, tmp_path):
    """decodeRecordedFile() for a batch: .iq / .c2 files (unnormalised amplitudes, a short recording, a missing file, a
    wrong extension) -> decoder_results identical to readRawIQfile's normalisation + ft8_subsystem on the CPU."""
    paths, want = [], []
    for k, (i_s, q_s) in enumerate(slots[:3]):
        n = 48000 if k != 1 else 41000
        gain = np.float32([0.02, 7.5, 1.0][k])
        inter = np.empty(2 * n, np.float32)
        inter[0::2] = i_s[:n] * gain; inter[1::2] = -(q_s[:n] * gain)
        if k == 2:
            p = tmp_path / f"s{k}.c2"
            with open(p, "wb") as f:
                f.write(b"000000_0000.c2".ljust(14, b"\0") + np.int32(2).tobytes() + np.float64(7.074).tobytes() + inter.tobytes())
        else:
            p = tmp_path / f"s{k}.iq"
            inter.tofile(p)
        paths.append(str(p))
        fi = np.zeros(48000, np.float32); fq = np.zeros(48000, np.float32)
        fi[:n] = inter[0::2]; fq[:n] = -inter[1::2]
        ci, cq, _ = oracle.condition(fi, fq, n)   # same expression as readRawIQfile's normalisation (rtlsdr_ft8d.c:762-778)
        want.append(oracle.subsystem(ci, cq))
    paths += [str(tmp_path / "missing.iq"), str(tmp_path / "wrong.txt")]
    res, nres, ns = pkg.decode_iq_files(ctx, paths)
    assert list(ns) == [48000, 41000, 48000, 0, 0] and nres[3] == 0 and nres[4] == 0
    for k in range(3):
        assert nres[k] == want[k]["n"] and res[k].tobytes() == want[k]["results"].tobytes()
    assert nres[0] >= 1


# ------------------------------------------------------------------------------------------- device-side synthesis
def _random_signals(pkg, rng, n, f_lo, f_hi, t_lo, t_hi, amp):
    items, texts = [], []
    for _ in range(n):
        to, de, ex = synth.random_message(rng)
        items.append((pkg.pack77_std(to, de, ex), float(rng.uniform(f_lo, f_hi)), float(rng.uniform(t_lo, t_hi)), amp))
        texts.append(f"{to} {de} {ex}")
    return pkg.make_signals(items), texts


def test_encode_tones_on_device(pkg, ctx, oracle):
    rng = np.random.default_rng(17)
    payloads = np.stack([np.frombuffer(oracle.pack_std(*synth.random_message(rng)), np.uint8) for _ in range(200)])
    t8 = ctx.encode_tones(payloads, 1)
    t4 = ctx.encode_tones(payloads, 0)
    for k in range(payloads.shape[0]):
        assert np.array_equal(t8[k], oracle.tones(payloads[k].tobytes()))
        assert np.array_equal(t4[k], ft8enc.tones_ft4(payloads[k].tobytes()))


def test_synth_matches_cpu_twin_and_decodes(pkg, ctx, oracle):
    """ft8b200_synth_raw / _slots / _audio == oracle/ft8_oracle_synth.c bit for bit (bytes, float bit patterns), for
    several slots with several signals each, and the noise stream depends only on (seed, slot index)."""
    rng = np.random.default_rng(23)
    # raw bytes: 3 slots x 400 000 samples, signals starting inside the window (incl. negative start)
    sigs, first = [], [0]
    for s in range(3):
        g, _ = _random_signals(pkg, rng, s + 1, 200.0, 1400.0, -0.05, 0.05, 20.0 + 10 * s)
        sigs.append(g); first.append(first[-1] + g.size)
    allsig = np.concatenate(sigs)
    n_samp = 400_000
    raw = ctx.synth_raw(allsig, first, 30.0, 99, first_slot_index=5, bytes_per_slot=2 * n_samp).cpu().numpy()
    for s in range(3):
        want = oracle.synth_raw(sigs[s], 30.0, 99, 5 + s, n_samp)
        assert np.array_equal(raw[s, :2 * n_samp], want), f"raw slot {s}"
    again = ctx.synth_raw(sigs[2], [0, sigs[2].size], 30.0, 99, first_slot_index=7, bytes_per_slot=2 * n_samp).cpu().numpy()
    assert np.array_equal(again[0, :2 * n_samp], raw[2, :2 * n_samp]), "slot content must not depend on its position in the batch"
    assert 25.0 < raw[0, 1::2].astype(np.float64).std() < 40.0   # ~30 LSB of noise + a 20 LSB tone
    # complex 3200 sps slots
    g, texts = _random_signals(pkg, rng, 4, 100.0, 1400.0, 0.2, 1.5, 0.3)
    d_i, d_q = ctx.synth_slots(np.concatenate([g, g[:1]]), [0, 4, 5], 1.0, 5)
    for s, part in enumerate((g, g[:1])):
        wi, wq = oracle.synth_float(1, False, part, 1.0, 5, s, 48000)
        assert bits_equal(d_i[s].cpu().numpy(), wi) and bits_equal(d_q[s].cpu().numpy(), wq)
    peak = torch.maximum(d_i.abs().amax(1), d_q.abs().amax(1))
    ctx.process_conditioned(d_i, d_q, peak)
    res, n = ctx.fetch_results(2)
    o = oracle.subsystem(*oracle.condition(*oracle.synth_float(1, False, g, 1.0, 5, 0, 48000), 48000)[:2])
    assert n[0] == o["n"] >= 3 and res[0].tobytes() == o["results"].tobytes()
    # 12 kHz audio, FT8 and FT4
    for proto, n_samp in ((1, 180_000), (0, 90_000)):
        g, texts = _random_signals(pkg, rng, 3, 300.0, 2500.0, 0.3, 1.0, 0.1)
        d_a = ctx.synth_audio(g, [0, 3], proto, 0.05, 11, n_samples=n_samp)
        wa, _ = oracle.synth_float(2, proto == 0, g, 0.05, 11, 0, n_samp)
        assert bits_equal(d_a[0].cpu().numpy(), wa)
        lines = [pkg.format_decoded(r) for r in pkg.decode_audio(ctx, d_a, 12000, proto)[0]]
        assert lines == oracle.decode_ft8_lines(wa, 12000, protocol=proto)
        assert {l.split("~  ")[1] for l in lines} == set(texts)


def test_synth_full_raw_slots_decode(pkg, ctx):
    """Full-size config #2 inputs made on the device: 4 raw 72 MB slots, one message each -> the whole path decodes them."""
    rng = np.random.default_rng(31)
    g, texts = _random_signals(pkg, rng, 4, 300.0, 1300.0, 0.3, 0.8, 20.0)
    for k in range(4):
        texts[k] = "CQ " + " ".join(texts[k].split()[1:2]) + " FN20"
        g[k]["payload"] = np.frombuffer(pkg.pack77_std(*texts[k].split()), np.uint8)
    raw = ctx.synth_raw(g, [0, 1, 2, 3, 4], 30.0, 1)
    ctx.process_raw(raw, 4)
    res, n = ctx.fetch_results(4)
    for k in range(4):
        assert n[k] >= 1 and res[k][0]["call"].decode() == texts[k].split()[1] and res[k][0]["loc"] == b"FN20"


# ------------------------------------------------------------------------------------------- GFSK (gen_ft8.c:28-102)
def _gfsk_signals(pkg, rng, n, f_lo, f_hi, t_lo, t_hi, amp_lo, amp_hi):
    items, texts = [], []
    for _ in range(n):
        to, de, ex = synth.random_message(rng)
        items.append((pkg.pack77_std(to, de, ex), float(rng.uniform(f_lo, f_hi)), float(rng.uniform(t_lo, t_hi)), float(rng.uniform(amp_lo, amp_hi))))
        texts.append(f"{to} {de} {ex}")
    return pkg.make_signals(items, gfsk=True), texts


def test_gfsk_synth_matches_cpu_twin(pkg, ctx, oracle):
    """GFSK mode of the synthesiser (Gaussian-smoothed frequency from an integer pulse table, extended end symbols, raised-cosine
    ramps) == its CPU twin bit for bit in all three output forms and both protocols; the twin is compared with the reference's own
    synth_gfsk() in tests/test_oracle_vs_ref.py.  Mixed FSK / GFSK signals in one slot are allowed."""
    rng = np.random.default_rng(61)
    n_samp = 3 * 384_000 + 5000                      # raw bytes: the ramp, three symbols and their pulse overlaps
    g, _ = _gfsk_signals(pkg, rng, 2, 300.0, 1300.0, -0.02, 0.01, 20.0, 35.0)
    g[1]["reserved"][0] = 0                          # one plain-FSK signal next to a GFSK one
    raw = ctx.synth_raw(g, [0, 2], 30.0, 5, first_slot_index=3, bytes_per_slot=2 * n_samp).cpu().numpy()
    assert np.array_equal(raw[0, :2 * n_samp], oracle.synth_raw(g, 30.0, 5, 3, n_samp))
    g, texts = _gfsk_signals(pkg, rng, 5, 100.0, 1400.0, 0.2, 1.5, 0.25, 0.4)
    d_i, d_q = ctx.synth_slots(g, [0, 5], 1.0, 8)
    wi, wq = oracle.synth_float(1, False, g, 1.0, 8, 0, 48000)
    assert bits_equal(d_i[0].cpu().numpy(), wi) and bits_equal(d_q[0].cpu().numpy(), wq)
    peak = torch.maximum(d_i.abs().amax(1), d_q.abs().amax(1))
    ctx.process_conditioned(d_i, d_q, peak)
    res, n = ctx.fetch_results(1)
    o = oracle.subsystem(*oracle.condition(wi, wq, 48000)[:2])
    assert n[0] == o["n"] >= 4 and res[0].tobytes() == o["results"].tobytes()
    for proto, n_samp in ((1, 180_000), (0, 90_000)):
        g, texts = _gfsk_signals(pkg, rng, 3, 300.0, 2500.0, 0.3, 1.0, 0.08, 0.12)
        d_a = ctx.synth_audio(g, [0, 3], proto, 0.05, 11, n_samples=n_samp)
        wa, _ = oracle.synth_float(2, proto == 0, g, 0.05, 11, 0, n_samp)
        assert bits_equal(d_a[0].cpu().numpy(), wa)
        lines = [pkg.format_decoded(r) for r in pkg.decode_audio(ctx, d_a, 12000, proto)[0]]
        assert lines == oracle.decode_ft8_lines(wa, 12000, protocol=proto)
        assert {l.split("~  ")[1] for l in lines} == set(texts)


def test_crowded_band_gfsk_parity(pkg, ctx500, oracle):
    """BASELINE config #3 with on-air-shaped signals: 60 overlapping GFSK messages over 50-1500 Hz at -24..+5 dB, K = 500 / 200
    messages: decoder_results identical to the oracle's ft8_subsystem on the twin's samples, for two slots."""
    def amp(snr_db):
        return float(np.sqrt(2.0 * (2500.0 / 3200.0) * 10.0 ** (snr_db / 10.0)))
    sigs, firsts = [], [0]
    for s in range(2):
        g, _ = _gfsk_signals(pkg, np.random.default_rng(900 + s), 60, 50.0, 1500.0, -0.5, 1.5, amp(-24.0), amp(5.0))
        sigs.append(g); firsts.append(firsts[-1] + g.size)
    d_i, d_q = ctx500.synth_slots(np.concatenate(sigs), firsts, 1.0, 17)
    peak = torch.maximum(d_i.abs().amax(1), d_q.abs().amax(1))
    ctx500.process_conditioned(d_i, d_q, peak)
    res, n = ctx500.fetch_results(2)
    total = 0
    for s in range(2):
        wi, wq = oracle.synth_float(1, False, sigs[s], 1.0, 17, s, 48000)
        assert bits_equal(d_i[s].cpu().numpy(), wi) and bits_equal(d_q[s].cpu().numpy(), wq)
        o = oracle.subsystem(*oracle.condition(wi, wq, 48000)[:2], max_cand=500, max_msgs=200)
        assert n[s] == o["n"] and res[s].tobytes() == o["results"].tobytes()
        total += int(n[s])
    assert total >= 20


def test_bulk_randomised_parity_sweep(pkg, ctx, ctx500, oracle):
    """Breadth rather than a new case: 64 slots of 1..60 overlapping GFSK signals (-24..+5 dB, random DT / frequency) through
    waterfall -> sync -> LDPC -> CRC -> unpack -> spot table in ONE batch, every slot's decoder_results[] (content, order, stale gaps)
    against the oracle's ft8_subsystem on the same samples -- with the daemon's limits (K = 120 / 50 messages) and with config #3's
    (K = 500 / 200); then 24 recordings of the 12 kHz path (FT8 and FT4) against decode_ft8's main()."""
    def amp(snr_db):
        return float(np.sqrt(2.0 * (2500.0 / 3200.0) * 10.0 ** (snr_db / 10.0)))
    rng = np.random.default_rng(4242)
    n_slots = 64
    sigs, firsts = [], [0]
    for s in range(n_slots):
        g, _ = _gfsk_signals(pkg, rng, 1 + (s * 59) // (n_slots - 1), 50.0, 1500.0, -0.5, 1.5, amp(-24.0), amp(5.0))
        sigs.append(g); firsts.append(firsts[-1] + g.size)
    d_i, d_q = ctx500.synth_slots(np.concatenate(sigs), firsts, 1.0, 23)
    peak = torch.maximum(d_i.abs().amax(1), d_q.abs().amax(1))
    h_i, h_q = d_i.cpu().numpy(), d_q.cpu().numpy()
    decoded = 0
    for c, K, M in ((ctx, 120, 50), (ctx500, 500, 200)):
        c.process_conditioned(d_i, d_q, peak)
        res, n = c.fetch_results(n_slots)
        for s in range(n_slots):
            o = oracle.subsystem(*oracle.condition(h_i[s], h_q[s], 48000)[:2], max_cand=K, max_msgs=M)
            assert n[s] == o["n"] and res[s].tobytes() == o["results"].tobytes(), (K, s)
            decoded += int(n[s])
    assert decoded >= 20 * n_slots // 4
    for proto, n_samp, n_rec in ((1, 180_000, 16), (0, 90_000, 8)):
        items, first = [], [0]
        for r in range(n_rec):
            for _ in range(1 + 4 * r):
                to, de, ex = synth.random_message(rng)
                items.append((pkg.pack77_std(to, de, ex), float(rng.uniform(200.0, 2900.0)), float(rng.uniform(0.0, 1.5)), float(rng.uniform(0.02, 0.5))))
            first.append(len(items))
        aud = ctx.synth_audio(pkg.make_signals(items, gfsk=True), first, proto, 0.05, 29, n_samples=n_samp)
        lines = pkg.decode_audio(ctx, aud, 12000, proto)
        h_a = aud.cpu().numpy()
        for r in range(n_rec):
            assert [pkg.format_decoded(x) for x in lines[r]] == oracle.decode_ft8_lines(h_a[r], 12000, protocol=proto), (proto, r)


def test_create_destroy_cycles_return_their_memory(pkg, raw_slot):
    """A daemon library is created and torn down many times over a process's life (initFFTW/freeFFTW per run, a pipe per band
    change): contexts, pipes with and without an SM partition, receiver streams and monitors are created, USED and destroyed
    repeatedly; the device's free memory afterwards is what it was after the first cycle (no cudaMalloc left behind)."""
    small = raw_slot[:12016 * 40]
    aud = np.zeros(1920 * 4, np.float32)

    def cycle(k):
        c = pkg.Context(0)
        d_raw = torch.from_numpy(small).to(dev())
        c.process_raw(d_raw, 1, small.size)
        c.fetch_results(1)
        st = pkg.Stream(c)
        st.callback(small[:65536]); st.flip(); st.fetch()
        st.close()
        c.close()
        p = pkg.Pipe(0, depth=3)
        if k % 2:
            try:
                p.set_partition(32)
            except pkg.Ft8Error:
                pass   # a driver without green contexts
        p.submit(d_raw, 1, small.size); p.collect(1)
        p.submit_host(small, 1, small.size); p.collect(1)
        p.close()
        m = pkg.Monitor()
        for o in range(0, aud.size, 1920):
            m.process(aud[o:o + 1920])
        m.close()
        del d_raw
        torch.cuda.synchronize()
        torch.cuda.empty_cache()
        return torch.cuda.mem_get_info()[0]

    cycle(0); cycle(1)              # first use: module load, per-device tables, the default context of the monitor
    free0 = cycle(2)
    for k in range(3, 15):
        free = cycle(k)
    assert free0 - free < (8 << 20), f"{(free0 - free) >> 20} MiB of device memory did not come back after 12 create/destroy cycles"


@pytest.mark.parametrize("dims", [dict(num_blocks=1, num_bins=16, time_osr=1, freq_osr=1), dict(num_blocks=3, num_bins=9, time_osr=2, freq_osr=1),
                                  dict(num_blocks=6, num_bins=8, time_osr=1, freq_osr=2), dict(num_blocks=13, num_bins=40, time_osr=2, freq_osr=2),
                                  dict(num_blocks=20, num_bins=7, time_osr=2, freq_osr=2), dict(num_blocks=5, num_bins=3, time_osr=1, freq_osr=1)])
def test_dropin_on_tiny_waterfalls(pkg, oracle, dims):
    """ft8_find_sync()/ft8_decode() on waterfalls far smaller than a slot -- a monitor that has seen a few blocks, a band of a few
    bins: every sync block hangs over an edge, most symbols of a candidate are out of range (zero LLRs), and below 8 bins the
    reference's loop never runs (decode.c:189): the same candidates, statuses and return values as the oracle, no error."""
    rng = np.random.default_rng(dims["num_blocks"] * 100 + dims["num_bins"])
    cells = dims["num_blocks"] * dims["time_osr"] * dims["freq_osr"] * dims["num_bins"]
    for proto in (1, 0):
        for t in (10, -50):
            mag = rng.integers(0, 256, cells, dtype=np.uint8)
            g = pkg.ft8_find_sync(mag, 30, t, protocol=proto, **dims)
            o = oracle.find_sync(mag, max_cand=30, min_score=t, protocol=proto, **dims)
            assert g.tobytes() == o.tobytes(), (dims, proto, t)
            for c in list(g[:4]) + [np.array([(5, -3, 0, 0, 0)], cand_dtype)[0]]:
                if dims["num_bins"] < 8:
                    break
                ok, msg, st = pkg.ft8_decode(mag, c, 20, protocol=proto, **dims)
                d = oracle.decode(mag, c, max_iters=20, protocol=proto, **dims)
                assert ok == bool(d["ok"]) and st.tobytes() == d["status"].tobytes(), (dims, proto, t, c)


def test_stream_overrun_and_large_calls(pkg, ctx, oracle):
    """A decoder that falls behind: more than two slots of samples arrive -- in 30 MB calls, far beyond librtlsdr's 64 KB -- before
    the buffer is flipped.  The reference keeps filtering and drops what does not fit (`if (idx < 48000)`, rtlsdr_ft8d.c:196-200):
    the buffer holds the first 48 000 outputs, the count stops there, and the filter state the NEXT slot starts from is that of
    every byte received."""
    rng = np.random.default_rng(31)
    st = pkg.Stream(ctx)
    o_state = oracle.new_decim()
    oi, oq = [], []
    for _ in range(5):
        chunk = rng.integers(0, 256, size=30_000_000, dtype=np.uint8)
        st.callback(chunk)
        a, b = oracle.decim_feed(o_state, chunk, 30_000_000 // 1502 + 2)
        oi.append(a); oq.append(b)
    oi = np.concatenate(oi); oq = np.concatenate(oq)
    assert oi.size > 2 * 48000 and st.count() == 48000
    st.flip()
    gi, gq, n = st.fetch()
    assert n == 48000 and bits_equal(gi, oi[:48000]) and bits_equal(gq, oq[:48000])
    tail = rng.integers(0, 256, size=10_000_000 + 8 * 3, dtype=np.uint8)
    st.callback(tail)
    a, b = oracle.decim_feed(o_state, tail, 7000)
    st.flip()
    gi, gq, n = st.fetch()
    assert n == a.size and bits_equal(gi[:n], a) and bits_equal(gq[:n], b) and not gi[n:].any()
    st.close()


@pytest.mark.parametrize("rate,tosr,fosr,proto", [(8000, 2, 2, 1), (6000, 2, 2, 1), (16000, 2, 2, 1), (24000, 1, 2, 1), (12000, 1, 1, 1), (12000, 4, 1, 1),
                                                  (12000, 2, 4, 1), (8000, 2, 2, 0), (16000, 2, 2, 0), (12000, 3, 2, 0)])
def test_monitor_other_rates_and_oversampling(ctx, oracle, rate, tosr, fosr, proto):
    """decode_ft8 takes the WAV file's own sample rate and compile-time oversampling factors (decode_ft8.c:16-17,111-150): frame sizes
    other than the two the kernel has compile-time stage plans for (3840 / 1152 points) go through its generic mixed-radix stage loop
    (2560, 1920, 5120, 768, 7680 ... points; radix order 4, 2, 3, 5 like kf_factor): waterfall bytes identical to the oracle's monitor."""
    rng = np.random.default_rng(rate + 10 * tosr + fosr)
    n = int(rate * (15.0 if proto == 1 else 7.5))
    t = np.arange(n, dtype=np.float64) / rate
    aud = (0.05 * rng.standard_normal(n) + 0.2 * np.sin(2 * np.pi * 0.11 * rate * t) + 0.02 * np.sin(2 * np.pi * 0.031 * rate * t)).astype(np.float32)
    ref, info, _ = oracle.monitor_waterfall(aud, rate, tosr, fosr, proto)
    mag, nb = ctx.monitor_waterfall(torch.from_numpy(np.stack([aud, aud[::-1].copy()])).to(dev()), rate, tosr, fosr, proto)
    assert nb == int(info[4]) > 0
    got = mag.cpu().numpy()
    assert np.array_equal(got[0][: ref.size], ref), f"{int((got[0][: ref.size] != ref).sum())} of {ref.size} cells differ"
    ref2, _, _ = oracle.monitor_waterfall(aud[::-1].copy(), rate, tosr, fosr, proto)
    assert np.array_equal(got[1][: ref2.size], ref2)


@pytest.mark.parametrize("rate", [11025, 22050, 44100, 48000, 96000])
def test_monitor_exotic_rates_are_exact_or_refused(pkg, ctx, oracle, rate):
    """Audio rates whose frames have prime factors beyond 5 (11 025 Hz: 3528 = 2^3 3^2 7^2; 22 050, 44 100 likewise) go through
    kiss_fft's GENERIC butterfly (kf_bfly_generic, kiss_fft.c:192-229), restated in the kernel's generic stage loop: identical bytes.
    A frame too large for one CTA's shared memory (96 kHz: 30 720 points) is refused with a reason -- the batched call either
    reproduces the oracle's bytes or returns an error, never a waterfall that is merely plausible."""
    rng = np.random.default_rng(rate)
    aud = (0.1 * rng.standard_normal(int(rate * 2.0))).astype(np.float32)
    try:
        mag, nb = ctx.monitor_waterfall(torch.from_numpy(aud[None]).to(dev()), rate, 2, 2, 1)
    except pkg.Ft8Error as e:
        assert rate > 48000, f"{rate} Hz must be supported: {e}"
        # refused: with a reason, and without leaving its CUDA error behind as the runtime's "last error" for the application's own
        # next call to trip over (torch checks it after every launch)
        assert "monitor" in str(e) or "ft8b200" in str(e)
        assert int((torch.zeros(4, device=dev()) + 1).sum().item()) == 4
        torch.cuda.synchronize()
        return
    ref, info, _ = oracle.monitor_waterfall(aud, rate, 2, 2, 1)
    assert nb == int(info[4]) and np.array_equal(mag.cpu().numpy()[0][: ref.size], ref)


def test_two_contexts_from_two_threads(pkg, raw_slot):
    """Receivers share a process (256 streams in BASELINE config #5): two contexts on one device, each driven from its own host
    thread at the same time -- raw batches through the whole path, 3200 sps slots, and a pipe next to them.  Every call's records
    equal those of the same call made alone (nothing process-global is shared unsynchronised: tables, work counters, scratch)."""
    import threading
    d_raw = torch.from_numpy(np.stack([raw_slot, np.roll(raw_slot, 8 * 999), np.roll(raw_slot, 8 * 5000)])).to(dev())
    ctxs = [pkg.Context(0), pkg.Context(0)]
    pipe = pkg.Pipe(0, depth=2)

    def work(k, out):
        c = ctxs[k]
        for rep in range(6):
            n = 1 + (rep + k) % 3
            c.process_raw(d_raw[:n], n)
            res, cnt = c.fetch_results(n)
            out.append((n, np.array(res).tobytes(), np.array(cnt).tolist()))

    def work_pipe(out):
        for rep in range(6):
            n = 1 + rep % 3
            pipe.submit(d_raw[:n], n)
            res, cnt = pipe.collect(n)
            out.append((n, np.array(res).tobytes(), np.array(cnt).tolist()))

    alone = [[], [], []]
    work(0, alone[0]); work(1, alone[1]); work_pipe(alone[2])
    assert all(cnt[0] >= 1 for _, _, cnt in alone[0])
    for _ in range(3):
        together = [[], [], []]
        th = [threading.Thread(target=work, args=(0, together[0])), threading.Thread(target=work, args=(1, together[1])),
              threading.Thread(target=work_pipe, args=(together[2],))]
        for t in th: t.start()
        for t in th: t.join()
        assert together == alone
    pipe.close()
    for c in ctxs: c.close()


def test_monitor_dropin_at_11025_hz(pkg, oracle):
    """monitor_init/monitor_process with a WAV file's own rate of 11 025 Hz (block 1764, 3528-point frames, radix-7 stages) and
    ft8_find_sync on the result: the drop-in's waterfall bytes and candidates are the oracle's."""
    rng = np.random.default_rng(11025)
    aud = (0.1 * rng.standard_normal(11025 * 15)).astype(np.float32)
    mon = pkg.Monitor(11025, 2, 2, 1)
    assert (mon.me.block_size, mon.me.nfft, mon.me.wf.num_bins) == (1764, 3528, 882)
    for o in range(0, aud.size - 1764 + 1, 1764):
        mon.process(aud[o:o + 1764])
    ref, info, _ = oracle.monitor_waterfall(aud, 11025, 2, 2, 1)
    got = mon.mag()
    assert mon.me.wf.num_blocks == int(info[4]) and np.array_equal(got[: ref.size], ref)
    dims = dict(num_blocks=int(info[4]), num_bins=882, time_osr=2, freq_osr=2)
    assert mon.find_sync(40, 5).tobytes() == oracle.find_sync(ref, max_cand=40, min_score=5, **dims).tobytes()
    mon.close()


def test_decode_with_candidates_from_nowhere(pkg, ctx, oracle, slots, bp_variant):
    """Candidates that did not come from ft8_find_sync -- frequency offsets beyond the band or negative, sub-offsets beyond the
    oversampling factors, time offsets far outside the slot: the reference would read outside the waterfall (it range-checks the
    block only, decode.c:275-279).  The library reads nothing outside the slot (this test is part of the memcheck pass): such
    candidates fail like dead ones, the valid candidates next to them in the same launch decode as always, and the drop-in call
    answers false instead of taking the process down."""
    real = oracle.waterfall(*slots[0])
    good = oracle.find_sync(real)[:3]
    wild = [(20, 0, 30000, 0, 0), (20, 0, -5, 0, 0), (20, 3, 249, 0, 0), (20, 3, 100, 7, 0), (20, 3, 100, 1, 200), (20, 30000, 10, 0, 0),
            (20, -30000, 10, 1, 1), (20, 32767, 32767, 255, 255), (20, -32768, -32768, 255, 255)]
    cands = np.zeros((1, ctx.K), cand_dtype)
    for q, c in enumerate(list(good) + wild):
        cands[0, q] = tuple(c)
    n = len(good) + len(wild)
    d_mag = torch.from_numpy(real[None]).to(dev())
    d_cand = torch.from_numpy(cands.view(np.uint8).reshape(1, ctx.K, 8)).to(dev())
    ok, stage, status, msg, plain, llr = ctx.decode(d_mag, d_cand, torch.tensor([n], dtype=torch.int32, device=dev()), want_plain=True, want_llr=True)
    torch.cuda.synchronize()
    for q, c in enumerate(good):
        d = oracle.decode(real, c)
        assert int(ok[0, q]) == d["ok"] and bits_equal(np.nan_to_num(llr[0, q].cpu().numpy()), np.nan_to_num(d["llr"]))
    assert not ok[0, len(good):n].any(), "a candidate outside the waterfall cannot decode"
    for c in wild[:5]:
        okd, _, _ = pkg.ft8_decode(real, np.array([c], cand_dtype)[0], 20)
        assert okd is False
    assert int((torch.zeros(4, device=dev()) + 1).sum().item()) == 4
