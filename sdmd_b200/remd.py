"""REMD replica scheduling and exchange (REMD.c / Server.c / REMDExec.c) on top of replica objects.

A *replica* is anything with ``run(t_until=...)``, ``potential_energy()``, ``set_temperature(T)`` and ``stats()`` —
``sdmd_b200.capi.Replica`` (GPU) in production; the tests also drive the CPU oracle through the same code to pin the
protocol against the reference's ``sServer``.

Protocol restated from the reference:
  * REMDRun (REMD.c:29-126): run until an executed event carries ``currenttime`` past ``rate``; E = CalPotenE();
    exchange; ``targetTemperature`` = new T; ``rate += ExchangeRate``.
  * server (Server.c:92-211): collect (E, {num, T}) from every replica in connection order, one sequential
    Metropolis sweep (``rearrange``), hand every replica its new {num, T}.  Uniform draws are
    ``rand() / (RAND_MAX / 1)`` from a generator seeded with the server's seed argument.
  * sREMD (REMDExec.c:79-138): temperatures in Kelvin × 0.0019872041, printed with ``%lf`` and re-parsed with atof.
"""
from __future__ import annotations

import numpy as np

from . import capi

BOLTZMANN = 0.0019872041  # REMDExec.c:19
RAND_MAX = 2147483647


def reduced_temperature(kelvin: float) -> float:
    return float("%lf" % (kelvin * BOLTZMANN))  # REMDExec.c:118 + Initialization.c:193


class ServerStream:
    """The sServer random stream: srand(seed) then RANDOMVALUE(0,1) per needed draw (Server.c:16-18,57-58)."""

    def __init__(self, seed: int, block: int = 4096):
        self.seed, self.block = int(seed), block
        self.raw = capi.glibc_rand(self.seed, block)
        self.pos = 0

    def peek(self, n: int) -> np.ndarray:
        while self.pos + n > len(self.raw):
            self.raw = capi.glibc_rand(self.seed, 2 * len(self.raw))
        return 0 + self.raw[self.pos:self.pos + n].astype(np.float64) / (float(RAND_MAX) / (1 - 0))

    def advance(self, n: int):
        self.pos += n


def exchange_sweep(epot, nums, temps, stream: ServerStream):
    """One collect→rearrange→distribute round; returns new (nums, temps)."""
    n = len(epot)
    u = stream.peek(2 * n + 2)
    new_nums, new_temps, used, _, _ = capi.remd_rearrange(epot, nums, temps, u)
    stream.advance(used)
    return new_nums, new_temps


class RemdEnsemble:
    """All replicas in one process (one GPU holding several replicas, or the test oracle)."""

    def __init__(self, replicas, temps_reduced, exchange_rate: float, server_seed: int, refill=None, batched=None):
        self.reps = list(replicas)
        self.batched = all(isinstance(r, capi.Replica) for r in self.reps) if batched is None else batched
        self.nums = list(range(1, len(self.reps) + 1))
        self.temps = [float(t) for t in temps_reduced]
        self.rate = float(exchange_rate)
        self.next_x = self.rate                      # rate = ((int)(currenttime / outputRate) + 1) * outputRate, REMD.c:32
        self.stream = ServerStream(server_seed)
        self.refill = refill
        self.log = [[] for _ in self.reps]           # (time, T, num) per exchange, what SaveRE writes (DataSave.c:687-700)
        for r, t in zip(self.reps, self.temps):
            r.set_temperature(t)

    def _step_batched(self):
        """All replicas are GPU contexts on one device: advance them concurrently (one thread block each)."""
        n = len(self.reps)
        pending = list(range(n))
        times = [0.0] * n
        while pending:
            if self.refill:
                for k in pending:
                    self.refill(k, self.reps[k])
            stats = capi.run_many([self.reps[k] for k in pending], self.next_x)
            nxt = []
            for k, st in zip(pending, stats):
                times[k] = st["currenttime"]
                if st["stop_reason"] == 3:       # SDMD_STOP_NEED_RNG
                    nxt.append(k)
            pending = nxt
        return capi.potential_energy_many(self.reps), times

    def step(self):
        """Advance every replica to the next exchange boundary, then exchange."""
        if self.batched:
            e, times = self._step_batched()
            self.nums, self.temps = exchange_sweep(e, self.nums, self.temps, self.stream)
            for k, r in enumerate(self.reps):
                r.set_temperature(self.temps[k])
                self.log[k].append((times[k], self.temps[k], self.nums[k]))
            self.next_x += self.rate
            return e
        e, times = [], []
        for k, r in enumerate(self.reps):
            while True:
                if self.refill:
                    self.refill(k, r)
                st, _ = r.run(t_until=self.next_x)
                if st["stop_reason"] != 3:           # SDMD_STOP_NEED_RNG: refill and continue
                    break
            times.append(st["currenttime"])
            e.append(r.potential_energy())
        self.nums, self.temps = exchange_sweep(e, self.nums, self.temps, self.stream)
        for k, r in enumerate(self.reps):
            r.set_temperature(self.temps[k])
            self.log[k].append((times[k], self.temps[k], self.nums[k]))
        self.next_x += self.rate
        return e

    def log_text(self, k: int) -> str:
        return "".join("%5.2f %8.4f %2i\n" % rec for rec in self.log[k])


def exchange_distributed(dist, rank: int, world: int, epot: float, num: int, T: float, stream: ServerStream):
    """Rank-parallel exchange with torch.distributed as transport (gloo on CPU; the GPU path uses the NCCL
    all-gather inside libsdmd_cuda.so): all-gather (E, T, num), run the identical sweep everywhere, keep own entry."""
    import torch
    mine = torch.tensor([epot, T, float(num)], dtype=torch.float64)
    allv = [torch.zeros(3, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allv, mine)
    e = [float(v[0]) for v in allv]
    temps = [float(v[1]) for v in allv]
    nums = [int(v[2]) for v in allv]
    nums, temps = exchange_sweep(e, nums, temps, stream)
    return nums[rank], temps[rank]
