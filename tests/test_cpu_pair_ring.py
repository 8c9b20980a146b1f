"""The buffer protocol of csrc/step_pair.cu (the 256x256 one-pass step), replayed on the CPU.

A CTA owns a ring of NS chunk buffers; chunk number g = 8 k + c (heatmap k, sweep step c) lives in buffer g mod NS.  The
gradient of the sweep steps c >= STAGE0 goes back through the buffer the chunk came in (STS + bulk store); such a buffer
takes its next load only when the bulk store has read it -- thread 0 checks that one heatmap later, after the sweeps -- while
the buffers of the steps c < STAGE0 take theirs right after the sweep.  The kernel states the resulting rule as a
static_assert (STAGE0 > 2 NCH - 1 - NS unless NS >= 2 NCH); this test replays the program order of one CTA and checks, for
the configurations the library builds and for the ones the rule forbids, that

  * a sweep never waits for a chunk whose load is issued only later in program order (the dead-lock the rule excludes),
  * a load never lands in a buffer whose data has not been swept into registers or whose staged gradient has not been read,
  * the gradient is staged into a buffer that still belongs to its own chunk.

(The kernel itself is checked on the GPU: tests/test_gpu_step.py, tests/test_gpu_round2.py, compute-sanitizer logs under
profiles/.)"""

import pytest

NCH = 8


class Deadlock(Exception):
    pass


def replay(ns, stage0, heatmaps):
    total = NCH * heatmaps
    issued = set()          # chunks whose load has been issued
    swept = set()           # chunks read into registers
    staged = {}             # buffer -> chunk whose gradient sits in it, bulk store not yet known to have read it
    owner = {}              # buffer -> chunk it holds / is loading

    def issue(g):
        if g >= total:
            return
        assert g not in issued
        buf = g % ns
        prev = owner.get(buf)
        if prev is not None:
            assert prev in swept, 'load %d would overwrite chunk %d before it was read' % (g, prev)
            assert buf not in staged, 'load %d would overwrite the staged gradient of chunk %d' % (g, staged[buf])
        owner[buf] = g
        issued.add(g)

    for g in range(ns):                                   # kernel prologue: thread 0 fills the ring
        issue(g)
    for k in range(heatmaps):
        for c in range(NCH):                              # load sweep: mbar_wait on every chunk of this heatmap
            g = NCH * k + c
            if g not in issued:
                raise Deadlock('heatmap %d waits for chunk %d, whose load is issued later in program order' % (k, c))
            assert owner[g % ns] == g
            swept.add(g)
        for c in range(stage0):                           # `if (tid < kEarly) issue(g0 + NS + tid)`
            issue(NCH * k + c + ns)
        if stage0 < NCH and k > 0:                        # thread 0, after the sums: bulk_wait_read(), then the next loads
            for c in range(stage0, NCH):
                buf = (NCH * (k - 1) + c) % ns
                assert staged.pop(buf) == NCH * (k - 1) + c
            for c in range(stage0, NCH):
                issue(NCH * (k - 1) + c + ns)
        for c in range(stage0, NCH):                      # backward: STS into the chunk's own buffer, one bulk store per chunk
            g = NCH * k + c
            buf = g % ns
            assert owner[buf] == g, 'the gradient of chunk %d would be staged into a buffer re-armed for chunk %d' % (g, owner[buf])
            assert buf not in staged
            staged[buf] = g
    return len(issued)


@pytest.mark.parametrize('ns,stage0', [(13, 3), (24, 0), (13, 8), (14, 2), (13, 5), (16, 0), (24, 3)])
def test_configurations_the_rule_allows_run_through(ns, stage0):
    for heatmaps in (1, 2, 3, 7, 40):
        assert replay(ns, stage0, heatmaps) == NCH * heatmaps          # every chunk loaded exactly once


@pytest.mark.parametrize('ns,stage0', [(13, 2), (13, 1), (13, 0), (14, 1), (15, 0), (9, 6)])
def test_configurations_the_rule_forbids_dead_lock(ns, stage0):
    assert not (stage0 == NCH or ns >= 2 * NCH or stage0 > 2 * NCH - 1 - ns)      # the kernel's static_assert would fire
    with pytest.raises(Deadlock):
        replay(ns, stage0, 4)


def test_the_static_assert_is_exactly_the_dead_lock_condition():
    for ns in range(NCH + 1, 3 * NCH):
        for stage0 in range(NCH + 1):
            allowed = stage0 == NCH or ns >= 2 * NCH or stage0 > 2 * NCH - 1 - ns
            try:
                replay(ns, stage0, 5)
                ran = True
            except Deadlock:
                ran = False
            assert ran == allowed, (ns, stage0)
