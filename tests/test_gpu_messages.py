"""GPU parity of the message end of the path: the device unpacker on every message type
and reject path of unpack.c:18-427, the same payloads on the air (synthesised, decoded, compared per candidate), and the
edge cases of the daemon's duplicate table / CQ filter (rtlsdr_ft8d.c:1487-1520) -- all through the C ABI against the CPU
oracle, which tests/test_oracle_vs_ref.py pins to the unmodified reference on the same generators."""
import numpy as np
import pytest

from oracle.pyoracle import cand_dtype, msg_dtype, result_dtype, status_dtype
from tools import synth
from test_gpu_parity import check_slot_against_oracle, dev, view

pytestmark = pytest.mark.gpu
torch = pytest.importorskip("torch")


def test_device_unpack77_every_message_type(ctx, oracle):
    """200 000 stratified payloads (tools/synth.py::payload_fuzz: free text, telemetry, standard with every token class
    -- DE/QRZ/CQ, CQ nnn, CQ aaaa, hashed, /R, /P, grids, RRR/RR73/73, R+-nn --, non-standard calls with every flip/rpt/cq,
    and everything the reference rejects) through the kernel-side unpack77: status and text identical to the oracle."""
    payloads, labels = synth.payload_fuzz(20261, 200_000)
    texts, status = ctx.unpack77_batch(payloads)
    seen = {}
    for k in range(payloads.shape[0]):
        rc, t = oracle.unpack77(bytes(payloads[k]) + b"\0\0")
        assert int(status[k]) == rc, (labels[k], bytes(payloads[k]).hex(), int(status[k]), rc)
        if rc >= 0:
            assert texts[k] == t, (labels[k], bytes(payloads[k]).hex(), texts[k], t)
        seen[(labels[k].split("_")[0], rc)] = seen.get((labels[k].split("_")[0], rc), 0) + 1
    # every class was exercised, including both reject codes of the standard type
    for key in [("free", 0), ("telemetry", 0), ("n3", -1), ("i3", -1), ("nonstd", 0), ("std", 0), ("std", -1), ("std", -2)]:
        assert seen.get(key, 0) > 1000, (key, seen)
    # spot checks of the forms the reference prints (unpack.c:40-75,150-206,296-340)
    flat = set(texts)
    for needle in ("CQ 000", "RR73", "<...>"):
        assert any(needle in t for t in flat), needle
    assert any(t.endswith("/R") or "/R " in t for t in flat) and any("/P" in t for t in flat)
    assert "" in flat   # an empty free text unpacks to the empty string (status 0)


@pytest.mark.parametrize("seed", [1, 2])
def test_every_message_type_on_the_air(pkg, ctx, oracle, bp_variant_all, seed):
    """The same kinds of payloads as signals: 24 slots x 6 signals synthesised on the device at high SNR (any 77-bit payload:
    ft8b200_signal_t.payload), through waterfall -> sync -> LLR/LDPC/CRC -> unpack -> spot table, compared PER CANDIDATE with the
    oracle on the same waterfall: ok flag, stage, decode_status_t (incl. unpack_status of rejected payloads), message_t,
    decoder_results[] with the gaps non-CQ messages leave, and the first-seen message log."""
    n_slots, per = 24, 6
    payloads, labels = synth.payload_fuzz(77 + seed, n_slots * per)
    rng = np.random.default_rng(seed)
    order = rng.permutation(n_slots * per)          # mix the types across slots
    items = []
    for s in range(n_slots):
        base = rng.permutation(per)
        for j in range(per):
            k = int(order[s * per + j])
            f0 = 120.0 + 220.0 * float(base[j]) + float(rng.uniform(0.0, 60.0))
            items.append((bytes(payloads[k]), f0, float(rng.uniform(0.2, 1.2)), 0.25))
    sig = pkg.make_signals(items)
    first = np.arange(0, n_slots * per + 1, per, dtype=np.int32)
    d_i, d_q = ctx.synth_slots(sig, first, 1.0, 4242 + seed)
    peak = torch.maximum(d_i.abs().amax(1), d_q.abs().amax(1))
    mag = ctx.waterfall(d_i, d_q, peak)
    cand, ncand = ctx.find_sync(mag)
    ok, stage, status, msg, plain, llr = ctx.decode(mag, cand, ncand, want_plain=True, want_llr=True)
    res, nres, umsg, ufreq, uscore = ctx.spots(cand, ncand, ok, msg)
    torch.cuda.synchronize()
    hm = mag.cpu().numpy()
    n_rejected = n_unique = 0
    got_texts = set()
    for s in range(n_slots):
        o = check_slot_against_oracle(oracle, ctx, hm[s], cand, ncand, ok, stage, status, msg, plain, llr, res, nres, umsg, ufreq, uscore, s, ctx.K, ctx.M)
        n_unique += o["n"]
        got_texts |= {m["text"].decode() for m in o["msgs"]}
        st = view(status[s], status_dtype)[: int(ncand[s])]
        n_rejected += int(((stage[s, : int(ncand[s])].cpu().numpy() == 3) & (st["unpack_status"] < 0)).sum())
    # what went on the air came back: the unpackable payloads as text, the others as unpack failures behind a good CRC
    want = {}
    for k in order:
        rc, t = oracle.unpack77(bytes(payloads[int(k)]) + b"\0\0")
        want[t if rc >= 0 else None] = want.get(t if rc >= 0 else None, 0) + 1
    n_bad = want.pop(None, 0)
    assert len(got_texts & set(want)) >= 0.9 * len(want), (len(got_texts & set(want)), len(want))
    assert n_rejected >= 0.8 * n_bad > 0
    assert n_unique >= 0.85 * (n_slots * per - n_bad)
    # end to end through the fused call: same records
    ctx.process_conditioned(d_i, d_q, peak)
    r2, n2 = ctx.fetch_results(n_slots)
    assert np.array_equal(n2, nres.cpu().numpy()) and r2.tobytes() == view(res, result_dtype).tobytes()


@pytest.fixture(params=[0, 1], ids=["bp_nodes", "bp_edges"])
def bp_variant_all(request, pkg):
    pkg.set_decode_variant(request.param)
    yield request.param
    pkg.set_decode_variant(0)


def _msg(text: str, h: int):
    m = np.zeros(1, msg_dtype)
    m[0]["text"] = text.encode()
    m[0]["hash"] = h
    return m[0]


def spots_cases():
    """Hand-made decode results for the a15 edge cases.  -> list of (name, cand[], ok[], msg[], max_msgs, defined_in_reference)."""
    rng = np.random.default_rng(15)
    cases = []

    def cands(n, scores=None):
        c = np.zeros(n, cand_dtype)
        c["score"] = scores if scores is not None else np.sort(rng.integers(10, 60, n))[::-1]
        c["time_offset"] = rng.integers(-12, 24, n)
        c["freq_offset"] = rng.integers(0, 249, n)
        c["time_sub"] = rng.integers(0, 2, n)
        c["freq_sub"] = rng.integers(0, 2, n)
        return c

    # 1. two-token CQ, CQ with a 3-token directed form, long tokens (truncated by %.12s / %.6s), non-CQ gaps, a "CQ"-prefixed call
    texts = ["CQ K1JT", "CQ DX R6WA LN32", "K1ABC W9XYZ -15", "CQ 123 PA9XYZ JO22", "CQ PJ4/K1ABC", "<...> PJ4/K1ABC RR73", "CQ TEST W9XYZ/R EN37",
             "CQ0ABC K1JT FN20", "CQ", "QRZ K1JT FN20", "TNX BOB 73 GL", "0123456789ABCDEF01", "CQ K1JT FN20", "CQ K1JT FN20"]
    m = np.array([_msg(t, 1000 + 7 * k) for k, t in enumerate(texts)])
    m[-1]["hash"] = m[-2]["hash"]                        # an exact duplicate (same hash, same text)
    cases.append(("tokens", cands(len(texts)), np.ones(len(texts), np.uint8), m, 50, True))
    # 2. hash clashes: same hash different text (both kept, linear probing), same text different hash (both kept), wrap-around at the table end
    texts = ["CQ AA1AA FN20", "CQ BB2BB FN21", "CQ AA1AA FN20", "CQ CC3CC FN22", "CQ DD4DD FN23", "CQ EE5EE FN24", "CQ BB2BB FN21"]
    hashes = [49, 49, 99, 49 + 50, 48, 49 + 100, 49]
    m = np.array([_msg(t, h) for t, h in zip(texts, hashes)])
    cases.append(("clashes", cands(len(texts)), np.ones(len(texts), np.uint8), m, 50, True))
    # 3. failed decodes and low scores are skipped; score below min_score with ok set must not reach the table
    texts = ["CQ K%dABC FN%02d" % (k % 10, k) for k in range(40)]
    c = cands(40, np.concatenate([np.sort(rng.integers(10, 60, 34))[::-1], [9, 9, 5, 0, -3, -20]]))
    okv = (rng.random(40) < 0.6).astype(np.uint8)
    okv[34:] = 1
    m = np.array([_msg(t, int(rng.integers(0, 1 << 14))) for t in texts])
    cases.append(("skips", c, okv, m, 50, True))
    # 4. exactly max_msgs - 1 unique messages + duplicates of them: the fullest table the reference survives with new messages still arriving
    texts = ["CQ N%dXY%s AA%02d" % (k % 10, chr(65 + k // 10), k) for k in range(49)]
    m = np.array([_msg(t, int(rng.integers(0, 1 << 14))) for t in texts] * 2)
    cases.append(("almost_full", cands(98), np.ones(98, np.uint8), m, 50, True))
    # 5. (undefined in the reference: it probes forever, rtlsdr_ft8d.c:1490-1502) 120 unique messages into 50 table slots: the 51st.. are dropped
    texts = ["CQ N%dXY%s AA%02d" % (k % 10, chr(65 + (k // 10) % 26), k % 100) + ("" if k < 100 else "X") for k in range(120)]
    m = np.array([_msg(t, int(rng.integers(0, 1 << 14))) for t in texts])
    cases.append(("overfull", cands(120), np.ones(120, np.uint8), m, 50, False))
    # 6. (undefined in the reference: strtok() returns NULL, :1509-1510) an empty text and an all-blank text are "not CQ"
    texts = ["", "CQ K1JT FN20", "   ", "CQ  W9XYZ   EN37 "]
    m = np.array([_msg(t, 33 + k) for k, t in enumerate(texts)])
    cases.append(("empty_text", cands(4), np.ones(4, np.uint8), m, 50, False))
    return cases


def test_spots_edge_cases(pkg, ctx, oracle):
    """ft8b200_spots on hand-made decode results: records, count and the first-seen log identical to the oracle's table logic,
    which test_oracle_vs_ref.py::test_spots_vs_reference_loop pins to the reference's own loop for the defined cases."""
    for name, c, okv, m, max_msgs, _ in spots_cases():
        assert max_msgs == ctx.M
        K = ctx.K
        assert c.size <= K
        cand = np.zeros((1, K), cand_dtype); cand[0, : c.size] = c
        ok = np.zeros((1, K), np.uint8); ok[0, : c.size] = okv
        msg = np.zeros((1, K), msg_dtype); msg[0, : c.size] = m
        d_cand = torch.from_numpy(cand.view(np.uint8).reshape(1, K, 8)).to(dev())
        d_msg = torch.from_numpy(msg.view(np.uint8).reshape(1, K, 28)).to(dev())
        res, nres, umsg, ufreq, uscore = ctx.spots(d_cand, torch.tensor([c.size], dtype=torch.int32, device=dev()), torch.from_numpy(ok).to(dev()), d_msg)
        torch.cuda.synchronize()
        o = oracle.spots(c, okv, m, max_msgs=max_msgs, min_score=10)
        assert int(nres[0]) == o["n"], name
        assert view(res[0], result_dtype).tobytes() == o["results"].tobytes(), name
        n = o["n"]
        gu = view(umsg[0], msg_dtype)[:n]
        assert np.array_equal(gu["text"], o["msgs"]["text"]) and np.array_equal(gu["hash"], o["msgs"]["hash"]), name
        assert np.array_equal(ufreq[0, :n].cpu().numpy().view(np.uint32), o["freq_hz"].view(np.uint32)) and np.array_equal(uscore[0, :n].cpu().numpy(), o["score"]), name
        if name == "overfull":
            assert n == 50
        if name == "tokens":
            r = view(res[0], result_dtype)
            assert r[0]["call"] == b"K1JT" and r[0]["loc"] == b"(null)" and r[1]["call"] == b"DX" and r[1]["loc"] == b"R6WA"
            assert r[8]["call"] == b"(null)" and r[2]["call"] == b""   # "CQ" alone; a non-CQ message leaves a gap
