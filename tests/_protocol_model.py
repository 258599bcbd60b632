"""A small exhaustive model of the launch / residency protocol of the multi-GPU step loop, one rank's view.
Test infrastructure only: it states the protocol's liveness argument in executable form and replays the three failures the hardware
showed in round 2; it models neither timing nor NVLink.

What is modelled
  * a GPU = `slots` CTA slots; a resident CTA holds `weight` slots until it exits; a CTA that waits, waits RESIDENT (spins);
  * main stream: step_1, step_2, ... (csrc/hp1.cu: hp1_step_kernel, `tiles` CTAs each).  Consecutive launches are chained by programmatic
    dependent launch: a kernel becomes launchable when every CTA of its predecessor has STARTED (the step executes
    griddepcontrol.launch_dependents right after claiming its tile); tile t of step T finishes only after tile t of step T-1 (per-tile
    done counter).  A *rare* step (no env has reset yet) additionally needs all of its tiles resident at once before any may finish
    (the one-sided grid barrier behind the stale-observation decision);
  * step T writes its observation into ring slot T % ring; push_T (csrc/p2p_allgather.cu, `push_ctas` CTAs on side stream T % ring)
    reads it after all tiles of step T have published, and step T+ring overwrites it: SAFETY = no tile of step T+ring finishes before
    push_T has finished;
  * three designs of the back-pressure / wake-up:
      "backpressure": the tiles of step T+ring spin in the kernel until push_T has finished reading; pushes spin for their producer;
      "gate"        : a one-CTA gate kernel in front of step T+ring waits for push_T (the step launches when the gate EXITS); pushes spin;
      "final"       : gate, and a one-CTA ready gate in front of every push waits for the producer (the push launches when it exits);
      "none"        : no protection at all (the checker must find the safety violation).
  * the scheduler is adversarial: whenever slots are free, ANY launchable kernel's next CTA may take them.

`explore()` walks every reachable state and returns (states, deadlocks, safety_violations)."""
from collections import deque


class Protocol:
    def __init__(self, *, steps, tiles, slots, ring, push_ctas, push_weight=1, design="final", rare=()):
        assert design in ("backpressure", "gate", "final", "none")
        self.steps, self.tiles, self.slots, self.ring = steps, tiles, slots, ring
        self.push_ctas, self.push_weight, self.design, self.rare = push_ctas, push_weight, design, frozenset(rare)
        # kernels: (kind, T); per-stream launch order
        main, side = [], {s: [] for s in range(ring)}
        for T in range(1, steps + 1):
            if design in ("gate", "final") and T > ring:
                main.append(("gate", T))
            main.append(("step", T))
            if design == "final":
                side[T % ring].append(("ready", T))
            side[T % ring].append(("push", T))
        self.kernels = main + [k for s in range(ring) for k in side[s]]
        self.index = {k: i for i, k in enumerate(self.kernels)}
        self.pred = {}
        for seq in [main] + list(side.values()):
            for a, b in zip(seq, seq[1:]):
                self.pred[b] = a

    # ---- state: per kernel, a tuple of per-CTA statuses 0 = pending, 1 = resident, 2 = done (non-step CTAs are interchangeable: sorted)
    def n_ctas(self, k):
        return self.tiles if k[0] == "step" else self.push_ctas if k[0] == "push" else 1

    def weight(self, k):
        return self.push_weight if k[0] == "push" else 1

    def initial(self):
        return tuple((0,) * self.n_ctas(k) for k in self.kernels)

    def st(self, state, k):
        return state[self.index[k]]

    def complete(self, state, k):
        return k not in self.index or all(c == 2 for c in self.st(state, k))

    def launchable(self, state, k):
        p = self.pred.get(k)
        if p is None:
            return True
        if p[0] == "step" and k[0] in ("step", "gate"):  # programmatic dependent launch: the predecessor's CTAs have all started
            return all(c != 0 for c in self.st(state, p))
        return self.complete(state, p)                    # gate -> step, ready -> push: on exit; plain stream order otherwise

    def published(self, state, T):
        return all(c == 2 for c in self.st(state, ("step", T)))

    def may_finish(self, state, k, i):
        kind, T = k
        if kind == "step":
            if T > 1 and self.st(state, ("step", T - 1))[i] != 2:
                return False                                            # per-tile chain
            if T in self.rare and any(c == 0 for c in self.st(state, k)):
                return False                                            # all tiles of a rare step must have arrived
            if self.design == "backpressure" and T > self.ring and not self.complete(state, ("push", T - self.ring)):
                return False
            return True
        if kind == "gate":
            return self.complete(state, ("push", T - self.ring))
        if kind == "ready":
            return self.published(state, T)
        return self.design == "final" or self.published(state, T)      # push: spins for its producer unless a ready gate did

    def used(self, state):
        return sum(self.weight(k) * sum(1 for c in s if c == 1) for k, s in zip(self.kernels, state))

    def successors(self, state):
        free = self.slots - self.used(state)
        out = []
        for ki, k in enumerate(self.kernels):
            s = state[ki]
            seen = set()
            for i, c in enumerate(s):
                key = (c, i) if k[0] == "step" else c   # tiles are distinguishable, other CTAs are not
                if key in seen:
                    continue
                seen.add(key)
                if c == 0 and self.weight(k) <= free and self.launchable(state, k):
                    out.append((ki, i, 1))
                elif c == 1 and self.may_finish(state, k, i):
                    out.append((ki, i, 2))
        res = []
        for ki, i, v in out:
            s = list(state[ki])
            s[i] = v
            if self.kernels[ki][0] != "step":
                s.sort(reverse=True)
            violation = None
            k = self.kernels[ki]
            if v == 2 and k[0] == "step" and k[1] > self.ring and not self.complete(state, ("push", k[1] - self.ring)):
                violation = f"tile {i} of step {k[1]} overwrote ring slot {k[1] % self.ring} before push {k[1] - self.ring} had read it"
            res.append((state[:ki] + (tuple(s),) + state[ki + 1:], violation))
        return res

    def explore(self, max_states=2_000_000):
        start = self.initial()
        seen, queue, deadlocks, violations = {start}, deque([start]), [], []
        final = tuple((2,) * self.n_ctas(k) for k in self.kernels)
        while queue:
            st = queue.popleft()
            succ = self.successors(st)
            if not succ and st != final:
                deadlocks.append(st)
            for nxt, bad in succ:
                if bad:
                    violations.append(bad)
                if nxt not in seen:
                    if len(seen) >= max_states:
                        raise RuntimeError("state space larger than max_states")
                    seen.add(nxt)
                    queue.append(nxt)
        assert final in seen or deadlocks, "the run can neither finish nor deadlock?"
        return len(seen), deadlocks, violations

    def describe(self, state):
        names = {0: "pending", 1: "RESIDENT", 2: "done"}
        return ", ".join(f"{k[0]}{k[1]}:" + "/".join(names[c] for c in s) for k, s in zip(self.kernels, state) if any(c == 1 for c in s))
