"""Two-level SMR cases (root Domain + one level-1 Domain, one Grid each): geometry from the golden fixture."""
import numpy as np

from cases import Case, VARS

NG = 4


class SmrCase(Case):
    def __init__(self, name):
        super().__init__(name)
        p = self.par
        self.build0 = self.build.replace("_smr", "")
        self.ion = self.build0 in ("ifront", "ioniz_sphere")
        self.nscal = 1 if self.ion else 0
        self.flux_i = p.par_getd("problem", "flux") if self.ion else 0.0
        self.has_grav = self.build0 == "ioniz_sphere"
        self.nx1 = tuple(p.par_geti("domain2", f"Nx{a}") for a in (1, 2, 3))
        self.disp1 = tuple(p.par_geti("domain2", k) for k in ("iDisp", "jDisp", "kDisp"))
        self.dx0 = tuple((self.xmax[a] - self.xmin[a]) / self.nx[a] for a in range(3))
        self.dx1 = tuple(self.dx0[a] / 2.0 for a in range(3))                      # init_mesh.c: dx = root dx / 2^level
        self.xmin1 = tuple(self.xmin[a] if self.disp1[a] == 0 else self.xmin[a] + float(self.disp1[a]) * self.dx1[a]
                           for a in range(3))                                       # init_mesh.c:377-382
        nv = 5 + self.nscal
        self.U0 = {v: self.gold[f"U0_{v}"] for v in VARS[:nv]}
        self.U1 = {v: self.gold[f"U1_{v}"] for v in VARS[:nv]}
        self.U0L1 = {v: self.gold[f"U0L1_{v}"] for v in VARS[:nv]}
        self.U1L1 = {v: self.gold[f"U1L1_{v}"] for v in VARS[:nv]}
        # child faces interior to the root get no physical BC (ProlongateLater, bvals_mhd.c:549-550)
        self.bc1 = []
        for a in range(3):
            self.bc1.append(self.bc[2 * a] if self.disp1[a] == 0 else 0)
            self.bc1.append(self.bc[2 * a + 1] if (self.disp1[a] + self.nx1[a]) // 2 == self.nx[a] else 0)
