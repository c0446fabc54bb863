"""Executable model of the work slicing of the streamed matvec (zgml_b200/csrc/qgemv_stream.cu): the launch is one linear space of
chunks, CTA c takes [c * per, (c + 1) * per), segments never straddle a block of column groups and are at most Lq chunks long,
and a cut column group's pieces are numbered in k order by a rule every CTA can evaluate on its own.  The kernel computes the
same quantities on the device; this model pins the invariants the merge relies on (every chunk exactly once, dense ordinals,
the host's bound on partial slots) for the plan parameters of the shapes that stream by default.  CPU only."""
import pytest


def plan(n_nb, n_kc, G, count, slots_max=296, per_min=16, Lq_records=128):
    """zg_qgemv_stream_plan with the default knobs: GB blocks of 8 column groups, nq chunks per group, `per` chunks per CTA."""
    GB, nq = -(-n_nb // 8), -(-n_kc // G)
    TQ, Lq = count * GB * nq, Lq_records // G
    per = max(per_min, -(-TQ // slots_max))
    if per < 2 * nq:
        if per < nq and TQ // nq >= 128:
            per = nq
        elif per < nq:
            while nq % per:
                per += 1
        else:
            per = -(-per // nq) * nq
    slots = -(-nq // Lq) + -(-nq // per) + 2
    return dict(GB=GB, nq=nq, TQ=TQ, Lq=Lq, per=per, grid=-(-TQ // per), slots=slots)


def segments(p):
    """What every CTA walks: (cta, linear block of column groups, first chunk inside it, chunks)."""
    for c in range(p["grid"]):
        q, hi = c * p["per"], min((c + 1) * p["per"], p["TQ"])
        while q < hi:
            gbl, qi = divmod(q, p["nq"])
            n = min(hi - q, p["nq"] - qi, p["Lq"])
            yield c, gbl, qi, n
            q += n


def ordinal_and_count(p, gbl, qi):
    """The device-side rule (no table): pieces of block `gbl` in k order, and the ordinal of the piece starting at chunk qi."""
    qq, end, cnt, ordinal = gbl * p["nq"], (gbl + 1) * p["nq"], 0, None
    q_mine = qq + qi
    while qq < end:
        hi = min((qq // p["per"] + 1) * p["per"], end)
        if qq <= q_mine < hi:
            ordinal = cnt + (q_mine - qq) // p["Lq"]
        cnt += -(-(hi - qq) // p["Lq"])
        qq = hi
    return ordinal, cnt


SHAPES = [  # (n_nb, n_kc, G, count): Llama-3-70B gate|up pair, down, LM head; its 2-way shard; the int4 microbench batches
    (896, 256, 4, 2), (256, 896, 4, 1), (4008, 256, 4, 1), (448, 256, 4, 2), (256, 448, 4, 1),
    (128, 128, 4, 8), (448, 128, 4, 8), (896, 256, 2, 2), (13, 33, 4, 3), (1, 1, 2, 1),
]


@pytest.mark.parametrize("n_nb,n_kc,G,count", SHAPES)
def test_every_chunk_once_and_dense_piece_ordinals(n_nb, n_kc, G, count):
    p = plan(n_nb, n_kc, G, count)
    assert p["grid"] <= 296 or p["per"] == 16            # one round of CTAs unless the floor of 16 chunks per CTA binds
    seen = {}
    pieces = {}
    for c, gbl, qi, n in segments(p):
        assert 1 <= n <= p["Lq"] and qi + n <= p["nq"]   # inside one block of column groups, inside the staging capacity
        for k in range(qi, qi + n):
            assert (gbl, k) not in seen
            seen[(gbl, k)] = c
        pieces.setdefault(gbl, []).append(qi)
    assert len(seen) == p["TQ"]                          # the whole linear space, nothing twice
    for gbl, starts in pieces.items():
        starts.sort()
        for idx, qi in enumerate(starts):
            ordinal, cnt = ordinal_and_count(p, gbl, qi)
            assert ordinal == idx and cnt == len(starts)  # every CTA derives the same dense numbering in k order
        assert len(starts) <= p["slots"]                 # the host's bound on partial slots per column group


def test_default_plans_of_the_70b_matvecs():
    # gate|up pair: whole column groups per warp (no cuts, no partials); down: 8 aligned pieces per group
    assert plan(896, 256, 4, 2)["per"] == 64 and plan(896, 256, 4, 2)["grid"] == 224
    d = plan(256, 896, 4, 1)
    assert d["nq"] % d["per"] == 0 and d["grid"] == 256
