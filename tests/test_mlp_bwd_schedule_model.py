"""CPU: a protocol model of the two pipelined schedules of the deformation-MLP backward (csrc/deform_mlp_bwd_tc5.cu, V2: two
alternating weight slots, one mbarrier; csrc/deform_mlp_bwd_tc5_db.cu: two dY images / weight slots / mbarriers).  The model
replays the kernels' control flow -- waits, cp.async weight copies, operand stores, MMA group commits, D_FE read-outs -- for
every head mask and several tile counts, with MMA groups completing as late as the waits allow, and checks the hazards the
schedules are built to avoid: no buffer is overwritten while a committed, not-yet-waited group still reads it, a weight image
is in its slot before the MMAs that read it, D_FE is stored before the next feature group overwrites it, and an mbarrier never
has two outstanding commits.  (The kernels themselves are checked on the GPU by tools/native/mlp_variant_check.)"""
import itertools


class Machine:
    """Buffers with 'pending readers' = committed MMA groups that have not been waited for."""
    def __init__(self):
        self.readers = {}            # buffer -> set of pending groups reading it
        self.content = {}            # buffer -> what it holds
        self.pending = {}            # group -> (barrier, buffers)
        self.outstanding = {}        # barrier -> group
        self.dfe = None              # tile whose d_feature sits in D_FE
        self.stored = []             # tiles whose d_feature reached global memory

    def write(self, buf, what):
        assert not self.readers.get(buf), f"{buf} overwritten while group(s) {self.readers[buf]} still read it"
        self.content[buf] = what

    def commit(self, group, barrier, reads, expect, writes_dfe=None):
        assert barrier not in self.outstanding, f"two outstanding commits on mbarrier {barrier}"
        for buf, what in expect.items():
            assert self.content.get(buf) == what, f"group {group}: {buf} holds {self.content.get(buf)}, wanted {what}"
        if writes_dfe is not None:
            assert self.dfe is None, f"D_FE of tile {self.dfe} overwritten before it was stored"
        self.pending[group] = (barrier, reads, writes_dfe)
        self.outstanding[barrier] = group
        for buf in reads:
            self.readers.setdefault(buf, set()).add(group)

    def wait(self, barrier):
        group = self.outstanding.pop(barrier, None)
        if group is None:
            return
        _, reads, writes_dfe = self.pending.pop(group)
        for buf in reads:
            self.readers[buf].discard(group)
        if writes_dfe is not None:
            self.dfe = writes_dfe

    def store_dfe(self):
        if self.dfe is not None:
            self.stored.append(self.dfe)
            self.dfe = None


def phases_of(mask):
    return [h for h in range(3) if (mask >> h) & 1] + [3]


def run_v2(mask, tiles):
    """deform_mlp_bwd_tc5_kernel<VER with bit 0>: group n reads weight slot n & 1; the NEXT group's image is copied after the
    drain of group n - 1; one mbarrier; operands (dY, X) single buffered."""
    m = Machine()
    order = [(t, ph) for t in range(tiles) for ph in phases_of(mask)]
    m.write("slot0", order[0][1])                                   # prologue: group 0's image
    for n, (t, ph) in enumerate(order):
        m.wait("bar")                                               # drain(): group n - 1
        m.store_dfe()
        if n + 1 < len(order):
            m.write(f"slot{(n + 1) & 1}", order[n + 1][1])          # the next group's image
        m.write("dY", (t, ph)); m.write("X", (t, "feat" if ph == 3 else "hidden")) if ph in (phases_of(mask)[0], 3) else None
        m.commit(n, "bar", {f"slot{n & 1}", "dY", "X"}, {f"slot{n & 1}": ph, "dY": (t, ph)}, writes_dfe=t if ph == 3 else None)
    m.wait("bar"); m.store_dfe()
    return m


def run_db(mask, tiles):
    """deform_mlp_bwd_tc5_db_kernel: group n uses dY image / weight slot / mbarrier n & 1; waits for group n - 2 at its start, for
    group n - 1 before X is rewritten (first head phase of a tile, feature phase)."""
    m = Machine()
    order = [(t, ph) for t in range(tiles) for ph in phases_of(mask)]
    fe_bar = None
    for n, (t, ph) in enumerate(order):
        b = n & 1

        def store_if_ready():
            if fe_bar is None or f"bar{fe_bar}" not in m.outstanding:
                m.store_dfe()
        if ph != 3:
            m.wait(f"bar{b}"); store_if_ready()
            m.write(f"slot{b}", ph)
            if ph == phases_of(mask)[0]:                            # X <- relu(hidden): the previous tile's feature group reads X
                m.wait(f"bar{b ^ 1}"); store_if_ready()
                m.write("X", (t, "hidden"))
            m.write(f"dY{b}", (t, ph))
            m.commit(n, f"bar{b}", {f"slot{b}", f"dY{b}", "X"}, {f"slot{b}": ph, f"dY{b}": (t, ph)})
        else:
            m.wait(f"bar{b}"); m.wait(f"bar{b ^ 1}"); store_if_ready()
            m.write(f"slot{b}", 3)
            m.write(f"dY{b}", (t, "scratch")); m.write(f"dY{b}", (t, 3)); m.write("X", (t, "feat"))
            fe_bar = b
            m.commit(n, f"bar{b}", {f"slot{b}", f"dY{b}", "X"}, {f"slot{b}": 3, f"dY{b}": (t, 3), "X": (t, "feat")}, writes_dfe=t)
    m.wait("bar0"); m.wait("bar1"); m.store_dfe()
    return m


def test_pipelined_schedules_have_no_buffer_hazards():
    for mask, tiles in itertools.product(range(8), (1, 2, 3, 5)):
        for run in (run_v2, run_db):
            m = run(mask, tiles)
            assert m.stored == list(range(tiles)), (run.__name__, mask, tiles, m.stored)
            assert not m.pending and not m.outstanding
