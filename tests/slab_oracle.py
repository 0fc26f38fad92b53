"""TEST INFRASTRUCTURE: a CPU backend for pibiti_b200.slab built on the oracle's stage functions, so
that the slab protocol (partition, migration, halo, boundary exchange) can be exercised without a GPU.
Same interface as pibiti_b200.slab.GpuSlabBackend; arrays are numpy."""
from __future__ import annotations

import numpy as np

from pibiti_b200.slab import REC, record_ids, z_cells


class OracleSlabBackend:
    def __init__(self, oracle, params, z_lo, z_hi, has_lower, has_upper):
        self.o = oracle
        self.par = np.ascontiguousarray(params).copy()
        self.z_lo, self.z_hi, self.has_lower, self.has_upper = z_lo, z_hi, has_lower, has_upper
        self.rec = np.zeros((0, REC), np.float32)
        self.ghosts = np.zeros((0, REC), np.float32)
        self.n_owned = 0

    def empty(self, width=REC):
        return np.zeros((0, width), np.float32)

    def set_params(self, params):
        self.par = np.ascontiguousarray(params).copy()

    def set_owned(self, records):
        self.rec = np.ascontiguousarray(np.asarray(records), np.float32).copy()
        self.n_owned = self.rec.shape[0]

    def get_owned(self):
        return self.rec.copy()

    def _zc(self, rec):
        return z_cells(rec[:, 0:4], self.par)

    def integrate(self):
        self.o.set_params(self.par)
        if self.rec.shape[0]:
            pos, vel = self.o.integrate(np.ascontiguousarray(self.rec[:, 0:4]), np.ascontiguousarray(self.rec[:, 4:8]))
            self.rec[:, 0:4], self.rec[:, 4:8] = pos, vel

    def take_leavers(self):
        zc = self._zc(self.rec)
        down = (zc < self.z_lo) & self.has_lower
        up = (zc >= self.z_hi) & self.has_upper
        out = self.rec[down].copy(), self.rec[up].copy()
        self.rec = self.rec[~(down | up)]
        return out

    def add_owned(self, recs):
        self.rec = np.concatenate([self.rec, np.asarray(recs, np.float32).reshape(-1, REC)], 0)

    def boundary_particles(self):
        zc = self._zc(self.rec)
        return (self.rec[(zc == self.z_lo) & self.has_lower].copy(), self.rec[(zc == self.z_hi - 1) & self.has_upper].copy())

    def add_ghosts(self, recs):
        self.ghosts = np.asarray(recs, np.float32).reshape(-1, REC).copy()

    def sort(self):
        allr = np.concatenate([self.rec, self.ghosts], 0)
        self.o.set_params(self.par)
        pos = np.ascontiguousarray(allr[:, 0:4])
        vel = np.ascontiguousarray(allr[:, 4:8])
        hashes = self.o.calc_hash(pos)[:, 0]
        ids = record_ids(allr)
        order = np.lexsort((ids, hashes)).astype(np.uint32)          # by cell hash, ties by ORIGINAL index
        pairs = np.ascontiguousarray(np.stack([hashes[order], order], 1).astype(np.uint32))
        ncell = int(self.par["numCells"][0])
        self.cell_start, self.spos, self.svel = self.o.reorder(pairs, pos, vel, ncell)
        self.sids = ids[order]
        self.szc = z_cells(self.spos, self.par)
        n = allr.shape[0]
        self.pairs = np.ascontiguousarray(np.stack([hashes[order], np.arange(n, dtype=np.uint32)], 1).astype(np.uint32))
        self.owned_mask = (self.szc >= self.z_lo) & (self.szc < self.z_hi)
        self.n_owned = int(self.owned_mask.sum())
        return int((self.szc < self.z_lo).sum()), self.n_owned, int((self.szc >= self.z_hi).sum())

    def density(self):
        self.pres, self.dens = self.o.density(self.spos, self.pairs, self.cell_start)

    def _dp_rows(self, mask):
        rows = np.zeros((int(mask.sum()), 8), np.float32)
        rows[:, 0:3] = self.spos[mask, 0:3]
        rows[:, 3] = self.pres[mask]
        rows[:, 4:7] = self.svel[mask, 0:3]
        rows[:, 7] = self.dens[mask]
        return rows

    def boundary_dp(self):
        return (self._dp_rows((self.szc == self.z_lo) & self.has_lower), self._dp_rows((self.szc == self.z_hi - 1) & self.has_upper))

    def set_ghost_dp(self, below, above):
        below, above = np.asarray(below, np.float32).reshape(-1, 8), np.asarray(above, np.float32).reshape(-1, 8)
        mb, ma = self.szc < self.z_lo, self.szc >= self.z_hi
        assert below.shape[0] == mb.sum() and above.shape[0] == ma.sum(), "ghost sets out of step between ranks"
        for rows, m in ((below, mb), (above, ma)):
            assert np.array_equal(rows[:, 0:3], self.spos[m, 0:3]), "ghost order differs from the owner's order"
            self.pres[m], self.dens[m] = rows[:, 3], rows[:, 7]

    def force(self):
        new_vel = self.o.force(self.spos, self.svel, self.pres, self.dens, self.pairs, self.cell_start)
        m = self.owned_mask
        rec = np.zeros((self.n_owned, REC), np.float32)
        rec[:, 0:4], rec[:, 4:8] = self.spos[m], new_vel[m]
        rec[:, 8] = self.sids[m].view(np.float32)
        rec[:, 9], rec[:, 10] = self.dens[m], self.pres[m]
        self.rec = rec
        self.ghosts = np.zeros((0, REC), np.float32)

    def sync(self):
        pass
