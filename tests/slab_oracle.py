"""TEST INFRASTRUCTURE: a CPU backend for pibiti_b200.slab built on the oracle's stage functions, so
that the slab protocol (partition, migration, halo, boundary exchange) can be exercised without a GPU.
Same interface and same message format as pibiti_b200.slab.GpuSlabBackend; arrays are numpy."""
from __future__ import annotations

import numpy as np

from pibiti_b200.slab import REC, record_ids, z_cells


class OracleSlabBackend:
    def __init__(self, oracle, params, z_lo, z_hi, has_lower, has_upper, caps):
        self.o = oracle
        self.par = np.ascontiguousarray(params).copy()
        self.z_lo, self.z_hi, self.has_lower, self.has_upper = z_lo, z_hi, has_lower, has_upper
        self.caps = caps
        self.rec = np.zeros((0, REC), np.float32)
        self.n_owned = 0

    # -- helpers ---------------------------------------------------------------------------------
    def _zc(self, rec):
        return z_cells(rec[:, 0:4], self.par)

    def _message(self, leavers, boundary):
        c = self.caps
        assert leavers.shape[0] <= c.leavers and boundary.shape[0] <= c.boundary, "message section overflow"
        m = np.zeros((c.rows, REC), np.float32)
        m[0, 0:2] = np.array([leavers.shape[0], boundary.shape[0]], np.uint32).view(np.float32)
        m[1:1 + leavers.shape[0]] = leavers
        m[1 + c.leavers:1 + c.leavers + boundary.shape[0]] = boundary
        return m

    def _sections(self, m):
        c = self.caps
        nl, nb = (int(v) for v in np.ascontiguousarray(m[0, 0:2]).view(np.uint32))
        return m[1:1 + nl], m[1 + c.leavers:1 + c.leavers + nb]

    def empty_message(self):
        return np.zeros((self.caps.rows, REC), np.float32)

    def empty_dp(self):
        return np.zeros((0, 8), np.float32)

    def sync(self):
        pass

    # -- interface -------------------------------------------------------------------------------
    def set_params(self, params):
        self.par = np.ascontiguousarray(params).copy()

    def set_owned(self, records):
        self.rec = np.ascontiguousarray(np.asarray(records), np.float32).copy()
        self.n_owned = self.rec.shape[0]

    def get_owned(self):
        return self.rec.copy()

    def integrate(self):
        self.o.set_params(self.par)
        if self.rec.shape[0]:
            pos, vel = self.o.integrate(np.ascontiguousarray(self.rec[:, 0:4]), np.ascontiguousarray(self.rec[:, 4:8]))
            self.rec[:, 0:4], self.rec[:, 4:8] = pos, vel

    def pack(self):
        zc = self._zc(self.rec)
        down = (zc < self.z_lo) & self.has_lower
        up = (zc >= self.z_hi) & self.has_upper
        self.left_down, self.left_up = self.rec[down].copy(), self.rec[up].copy()
        self.rec = self.rec[~(down | up)]
        zc = self._zc(self.rec)
        bnd_down = self.rec[(zc == self.z_lo) & self.has_lower]
        bnd_up = self.rec[(zc == self.z_hi - 1) & self.has_upper]
        return self._message(self.left_down, bnd_down), self._message(self.left_up, bnd_up)

    def unpack(self, below, above):
        lb, gb = self._sections(np.asarray(below, np.float32))
        la, ga = self._sections(np.asarray(above, np.float32))
        self.rec = np.concatenate([self.rec, lb, la], 0)
        # ghosts: the neighbours' boundary layers plus my own leavers (they now sit in those layers)
        self.ghosts = np.concatenate([gb, ga, self.left_down, self.left_up], 0)

    def sort(self):
        allr = np.concatenate([self.rec, self.ghosts], 0)
        zc_all = self._zc(allr)
        lo = self.z_lo - (1 if self.has_lower else 0)
        hi = self.z_hi + (1 if self.has_upper else 0)
        allr = allr[(zc_all >= lo) & (zc_all < hi)]              # outside the local table: dropped (dummy cell)
        self.o.set_params(self.par)
        pos = np.ascontiguousarray(allr[:, 0:4])
        vel = np.ascontiguousarray(allr[:, 4:8])
        hashes = self.o.calc_hash(pos)[:, 0]
        ids = record_ids(allr)
        order = np.lexsort((ids, hashes)).astype(np.uint32)       # by cell hash, ties by ORIGINAL index
        pairs = np.ascontiguousarray(np.stack([hashes[order], order], 1).astype(np.uint32))
        ncell = int(self.par["numCells"][0])
        self.cell_start, self.spos, self.svel = self.o.reorder(pairs, pos, vel, ncell)
        self.sids = ids[order]
        self.szc = z_cells(self.spos, self.par)
        n = allr.shape[0]
        self.pairs = np.ascontiguousarray(np.stack([hashes[order], np.arange(n, dtype=np.uint32)], 1).astype(np.uint32))
        self.owned_mask = (self.szc >= self.z_lo) & (self.szc < self.z_hi)
        self.n_owned = int(self.owned_mask.sum())
        self.ghost_counts = (int((self.szc < self.z_lo).sum()), int((self.szc >= self.z_hi).sum()))
        return self.ghost_counts[0], self.n_owned, self.ghost_counts[1]

    def density(self):
        self.pres, self.dens = self.o.density(self.spos, self.pairs, self.cell_start)

    def _dp_rows(self, mask):
        rows = np.zeros((int(mask.sum()), 8), np.float32)
        rows[:, 0:3] = self.spos[mask, 0:3]
        rows[:, 3] = self.pres[mask]
        rows[:, 4:7] = self.svel[mask, 0:3]
        rows[:, 7] = self.dens[mask]
        return rows

    def pack_dp(self):
        return (self._dp_rows((self.szc == self.z_lo) & self.has_lower), self._dp_rows((self.szc == self.z_hi - 1) & self.has_upper))

    def expected_dp(self):
        return self.ghost_counts

    def unpack_dp(self, below, above):
        below, above = np.asarray(below, np.float32).reshape(-1, 8), np.asarray(above, np.float32).reshape(-1, 8)
        mb, ma = self.szc < self.z_lo, self.szc >= self.z_hi
        assert below.shape[0] == mb.sum() and above.shape[0] == ma.sum(), "ghost sets out of step between ranks"
        for rows, m in ((below, mb), (above, ma)):
            assert np.array_equal(rows[:, 0:3], self.spos[m, 0:3]), "ghost order differs from the owner's order"
            self.pres[m], self.dens[m] = rows[:, 3], rows[:, 7]

    def force_interior(self):
        pass                            # the CPU backend evaluates everything once the ghost rows are in

    def force_boundary(self):
        self.force()

    def force(self):
        new_vel = self.o.force(self.spos, self.svel, self.pres, self.dens, self.pairs, self.cell_start)
        m = self.owned_mask
        rec = np.zeros((self.n_owned, REC), np.float32)
        rec[:, 0:4], rec[:, 4:8] = self.spos[m], new_vel[m]
        rec[:, 8] = self.sids[m].view(np.float32)
        rec[:, 9], rec[:, 10] = self.dens[m], self.pres[m]
        self.rec = rec
