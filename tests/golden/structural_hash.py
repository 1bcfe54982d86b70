"""Structural (template-level) hash of a recorded circuit -- for circuits too large to flatten (TEST INFRASTRUCTURE).

The flat canonical stream of emission_model.py needs memory proportional to the gate count; the pairing-level
compositions (final exponentiation 3.5 G gates, Miller loop 6.9 G) do not fit.  The same information -- every gate,
its type, its wiring, which outputs are dead, and the order of everything -- is also determined by the component
DAG: a component body is a sequence of gates and calls, and what a call contributes is fixed by the callee's body
and by which of its outputs the caller keeps alive.  This module hashes that DAG bottom-up:

    H(body, mask) = sha256( n_in
                            | for every run of gates:  "G", types, a, b, c      (c = DEAD when the output is dead)
                            | for every call:          "C", H(callee, callee mask), input wires, "O", output wires
                            | "E", the body's output wires )

with wires numbered canonically inside the body: 0 / 1 constants, 2.. inputs, then every live gate output and every
call output that is a new wire, in order of appearance.  `mask` says which of the body's outputs the caller uses; a
wire is live iff a gate of the body reads it, it is passed to a call, or it is an output the mask keeps (the
reference's credit rule, component_meta.rs / streaming_mode.rs).

Two independent producers are compared:
  * `model_hash(root_fn, n_inputs)`: the gadget restatement of emission_model.py run over `SCtx`, which records
    bodies once per component key and evaluates liveness per mask itself;
  * `product_hash(program)`: the product's recorder output (`Program.export_templates()`: its templates are already
    per (key, liveness mask), with dead outputs marked by its credit stacks).
Equal hashes mean equal flattened streams (gate order, types, wiring, dead gates), whatever the size.
"""
from __future__ import annotations

import hashlib

import numpy as np

DEAD = 0xFFFFFFFF


def _u32(x):
    return np.asarray(x, np.int64).astype(np.uint32).tobytes()


# ======================================================================================== model side
class _Body:
    __slots__ = ("n_in", "n_local", "items", "outs", "variants")


class SCtx:
    """Records one component body: gates in runs, calls as references to other bodies (never flattened)."""

    def __init__(self, n_inputs, memo=None):
        self.n_in = n_inputs
        self.next = 2 + n_inputs
        self.items = []                         # ("g", t, a, b, c) arrays | ("c", body, inputs, outs)
        self.t, self.a, self.b, self.c = [], [], [], []
        self.memo = {} if memo is None else memo

    def issue(self):
        w = self.next
        self.next += 1
        return w

    def gate(self, typ, a, b, c):
        self.t.append(typ)
        self.a.append(a)
        self.b.append(b)
        self.c.append(c)

    def _flush(self):
        if self.t:
            self.items.append(("g", np.array(self.t, np.uint8), np.array(self.a, np.int64), np.array(self.b, np.int64),
                               np.array(self.c, np.int64)))
            self.t, self.a, self.b, self.c = [], [], [], []

    def component(self, key, inputs, body):
        inputs = [int(w) for w in inputs]
        k = (key, len(inputs))
        rec = self.memo.get(k)
        if rec is None:
            sub = SCtx(len(inputs), self.memo)
            outs = body(sub, list(range(2, 2 + len(inputs))))
            rec = sub.close(outs)
            self.memo[k] = rec
        self._flush()
        # the caller's view of the outputs: constants and pass-through inputs are aliases, the rest are new wires
        # (one per distinct callee wire)
        outs, seen = [], {}
        for o in rec.outs:
            if o < 2:
                outs.append(o)
            elif o < 2 + rec.n_in:
                outs.append(inputs[o - 2])
            else:
                if o not in seen:
                    seen[o] = self.issue()
                outs.append(seen[o])
        self.items.append(("c", rec, inputs, list(outs)))
        return outs

    def close(self, outs):
        self._flush()
        rec = _Body()
        rec.n_in, rec.n_local, rec.items, rec.outs, rec.variants = self.n_in, self.next, self.items, [int(o) for o in outs], {}
        return rec


def _variant(rec, mask):
    """(hash, live flag per output position) of a body whose caller keeps the outputs in `mask` alive."""
    key = bytes(mask)
    got = rec.variants.get(key)
    if got is not None:
        return got
    live = np.zeros(rec.n_local, bool)
    for it in rec.items:
        if it[0] == "g":
            live[it[2]] = True
            live[it[3]] = True
        else:
            live[np.array(it[2], np.int64)] = True
    for o, m in zip(rec.outs, mask):
        if m and o >= 2:
            live[o] = True
    canon = np.full(rec.n_local, -1, np.int64)
    canon[:2 + rec.n_in] = np.arange(2 + rec.n_in)
    nxt = 2 + rec.n_in
    h = hashlib.sha256()
    h.update(b"T" + int(rec.n_in).to_bytes(4, "little"))
    for it in rec.items:
        if it[0] == "g":
            _, t, a, b, c = it
            lv = live[c]
            ids = nxt + np.cumsum(lv) - 1
            canon[c[lv]] = ids[lv]
            nxt += int(lv.sum())
            cc = np.where(lv, ids, DEAD)
            h.update(b"G")
            h.update(t.tobytes())
            h.update(_u32(canon[a]))
            h.update(_u32(canon[b]))
            h.update(_u32(cc))
        else:
            _, child, inputs, outs = it
            # which of the callee's outputs this body keeps alive (bits at alias positions change nothing in the callee)
            cmask = [1 if (co >= 2 + child.n_in and live[o]) else 0 for co, o in zip(child.outs, outs)]
            ch, child_live = _variant(child, cmask)
            ow = []
            for j, (co, o) in enumerate(zip(child.outs, outs)):
                if co < 2 + child.n_in:
                    ow.append(int(canon[o]) if o >= 2 else o)          # constant / pass-through: an existing wire
                elif not child_live[j]:
                    ow.append(DEAD)
                else:
                    if canon[o] < 0:
                        canon[o] = nxt
                        nxt += 1
                    ow.append(int(canon[o]))
            h.update(b"C" + ch)
            h.update(_u32(canon[np.array(inputs, np.int64)]))
            h.update(b"O")
            h.update(_u32(ow))
    outs_live = [bool(o < 2 + rec.n_in or live[o]) for o in rec.outs]
    h.update(b"E")
    h.update(_u32([int(canon[o]) if (o < 2 + rec.n_in or live[o]) else DEAD for o in rec.outs]))
    got = (h.digest(), outs_live)
    rec.variants[key] = got
    return got


def model_hash(root_fn, n_inputs):
    """root_fn(ctx, input_wires) -> output wires, written against the emission_model gadget API."""
    x = SCtx(n_inputs)
    rec = x.close(root_fn(x, list(range(2, 2 + n_inputs))))
    return _variant(rec, [1] * len(rec.outs))[0].hex()


# ======================================================================================== product side
def product_hash(program):
    """The same hash from the product's exported template DAG (gsv_program_export_templates)."""
    return templates_hash(*program.export_templates())


def templates_hash(root, tmpl, gates, calls, items, call_wires, outs):
    """Hash of an exported template DAG (the arrays of gsv_program_export_templates, from any recorder that went
    through the C ABI -- the product's own generator, or the reference's Rust gadgets behind gsv-cuda's GpuRecorder)."""
    tmpl = np.asarray(tmpl, np.int64).reshape(-1, 12)
    gates = np.asarray(gates, np.int64).reshape(-1, 4)
    calls = np.asarray(calls, np.int64).reshape(-1, 3)
    items = np.asarray(items, np.int64)
    call_wires = np.asarray(call_wires, np.int64)
    outs = np.asarray(outs, np.int64)
    done = {}

    def visit(ti):
        if ti in done:
            return done[ti]
        n_in, n_wires, g0, ng, c0, nc, i0, ni, w0, nw, o0, no = (int(v) for v in tmpl[ti])
        its = items[i0:i0 + ni]
        is_call = (its >> 31) & 1
        idx = its & 0x7FFFFFFF
        canon = np.full(n_wires + 1, -1, np.int64)
        canon[:2 + n_in] = np.arange(2 + n_in)
        nxt = 2 + n_in
        h = hashlib.sha256()
        h.update(b"T" + n_in.to_bytes(4, "little"))
        # runs of consecutive gate items
        k = 0
        n_items = len(its)
        call_pos = np.flatnonzero(is_call)
        bounds = [-1] + call_pos.tolist() + [n_items]
        for r in range(len(bounds) - 1):
            lo, hi = bounds[r] + 1, bounds[r + 1]
            if hi > lo:
                g = gates[g0 + idx[lo:hi]]
                a, b, c, t = g[:, 0], g[:, 1], g[:, 2], g[:, 3]
                lv = c != DEAD
                ids = nxt + np.cumsum(lv) - 1
                # a gate may read the output of an earlier gate of the same run: assign ids first (wires are SSA in the
                # recorded gadgets; an in-place overwrite would need the sequential form)
                canon[c[lv]] = ids[lv]
                nxt += int(lv.sum())
                h.update(b"G")
                h.update(t.astype(np.uint8).tobytes())
                h.update(_u32(canon[a]))
                h.update(_u32(canon[b]))
                h.update(_u32(np.where(lv, ids, DEAD)))
            if hi < n_items:
                callee, in_off, out_off = (int(v) for v in calls[c0 + idx[hi]])
                ch = visit(callee)
                c_n_in, c_no = int(tmpl[callee][0]), int(tmpl[callee][11])
                ins = call_wires[w0 + in_off:w0 + in_off + c_n_in]
                ow_local = call_wires[w0 + out_off:w0 + out_off + c_no]
                ow = []
                for o in ow_local.tolist():
                    if o == DEAD or o < 2:
                        ow.append(o)
                    else:
                        if canon[o] < 0:
                            canon[o] = nxt
                            nxt += 1
                        ow.append(int(canon[o]))
                h.update(b"C" + ch)
                h.update(_u32(np.where(ins == DEAD, DEAD, canon[np.minimum(ins, n_wires)])))
                h.update(b"O")
                h.update(_u32(ow))
        o_l = outs[o0:o0 + no]
        h.update(b"E")
        h.update(_u32([DEAD if o == DEAD else (o if o < 2 else int(canon[o])) for o in o_l.tolist()]))
        done[ti] = h.digest()
        return done[ti]

    import sys
    sys.setrecursionlimit(10000)
    return visit(int(root)).hex()
