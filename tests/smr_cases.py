"""SMR cases: a nest of single-Grid Domains (root + one Domain per finer level); geometry from the golden fixture.
`levels[n]` describes level n; the two-level attribute names (nx1, disp1, U0L1 ...) stay for the level-1 Domain."""
import numpy as np

from cases import Case, VARS

NG = 4


class SmrCase(Case):
    def __init__(self, name):
        super().__init__(name)
        p = self.par
        self.build0 = self.build.replace("_smr", "").replace("_mpi", "")
        self.ion = self.build0 in ("ifront", "ioniz_sphere")
        self.nscal = 1 if self.ion else 0
        self.flux_i = p.par_getd("problem", "flux") if self.ion else 0.0
        self.has_grav = self.build0 == "ioniz_sphere"
        self.nlevels = p.par_geti("job", "num_domains")
        self.dx0 = tuple((self.xmax[a] - self.xmin[a]) / self.nx[a] for a in range(3))
        nv = 5 + self.nscal
        self.levels = []
        for n in range(self.nlevels):
            blk = f"domain{n + 1}"
            nx = tuple(p.par_geti(blk, f"Nx{a}") for a in (1, 2, 3))
            disp = (0, 0, 0) if n == 0 else tuple(p.par_geti(blk, k) for k in ("iDisp", "jDisp", "kDisp"))
            dx = tuple(self.dx0[a] / 2.0 ** n for a in range(3))                    # init_mesh.c: dx = root dx / 2^level
            xmin = tuple(self.xmin[a] if disp[a] == 0 else self.xmin[a] + float(disp[a]) * dx[a]
                         for a in range(3))                                         # init_mesh.c:377-382
            bc = []        # faces interior to the root get no physical BC (ProlongateLater, bvals_mhd.c:549-550)
            for a in range(3):
                bc.append(self.bc[2 * a] if disp[a] == 0 else 0)
                bc.append(self.bc[2 * a + 1] if (disp[a] + nx[a]) // 2 ** n == self.nx[a] else 0)
            pre = "" if n == 0 else f"L{n}"
            self.levels.append(dict(
                nx=nx, disp=disp, dx=dx, xmin=xmin, bc=bc,
                U0={v: self.gold[f"U0{pre}_{v}"] for v in VARS[:nv]},
                U1={v: self.gold[f"U1{pre}_{v}"] for v in VARS[:nv]},
                edgeflux1=self.gold[f"U1{pre}_edgeflux"] if self.ion else None))
        l1 = self.levels[1]
        self.nx1, self.disp1, self.dx1, self.xmin1, self.bc1 = l1["nx"], l1["disp"], l1["dx"], l1["xmin"], l1["bc"]
        self.U0, self.U1 = self.levels[0]["U0"], self.levels[0]["U1"]
        self.U0L1, self.U1L1 = l1["U0"], l1["U1"]
