"""Total energy of a tight-binding model (reference: pysktb/forces.py:367-425).

Only `get_total_energy` is provided: the band energy runs on the device (mesh eigenvalues through the C ABI, reduced
by `sktb_band_energy`), the repulsive pair energy is summed on the host exactly as the reference does.  The
reference's `get_forces` is broken as shipped and is not reproduced.
"""
import numpy as np

from . import _lib


class Forces:
    def __init__(self, hamiltonian):
        self.ham = hamiltonian
        self.structure = hamiltonian.structure
        self.system = hamiltonian.system
        self.n_atoms = len(self.structure.atoms)
        self.n_orbitals = hamiltonian.n_orbitals

    def _get_kpoint_mesh(self, nk):
        """forces.py:76-81: linspace(endpoint=False) per axis (i * (1/n), not i/n - SURVEY Q10), kx outer."""
        kx, ky, kz = (np.linspace(0, 1, int(n), endpoint=False) for n in nk)
        return np.stack(np.meshgrid(kx, ky, kz, indexing="ij"), axis=-1).reshape(-1, 3)

    def get_total_energy(self, n_electrons, nk=[10, 10, 10], soc=True):
        """-> (E_total, E_band, E_rep).  soc=True fills n_electrons states, soc=False n_electrons // 2 with a
        factor 2 for spin (forces.py:381-395)."""
        import torch
        ham = self.ham
        kpts = self._get_kpoint_mesh(nk)
        n_k = len(kpts)
        soc = bool(soc)
        N = self.n_orbitals * 2 if soc else self.n_orbitals
        n_occ = min(int(n_electrons) if soc else int(n_electrons) // 2, N)
        with torch.cuda.device(ham._device):
            kd = torch.from_numpy(np.ascontiguousarray(kpts)).cuda()
            ev = torch.empty((N, n_k), dtype=torch.float64, device="cuda")
            gate = ham.asym_bound > 0.5e-9
            fl = torch.zeros(n_k, dtype=torch.int32, device="cuda") if gate else None
            lib = _lib.lib()
            _lib.check(lib.sktb_solve(ham._plan, kd.data_ptr(), n_k, int(soc), ev.data_ptr(), None, n_k, 0,
                                      fl.data_ptr() if gate else None, ham._stream()))
            out = torch.zeros(1, dtype=torch.float64, device="cuda")
            _lib.check(lib.sktb_band_energy(ham._plan, ev.data_ptr(), n_k, n_k, n_occ, out.data_ptr(), ham._stream()))
            s = float(out.cpu()[0])
        e_band = s / n_k * (1.0 if soc else 2.0)
        e_rep = self._get_repulsive_energy()
        return e_band + e_rep, e_band, e_rep

    def _get_repulsive_energy(self):
        """forces.py:407-425 + system.py:244-261: sum over bonded pairs i < j (all images) of params[pair]['repulsive'](d)."""
        e = 0.0
        st, sy = self.structure, self.system
        for i in range(self.n_atoms):
            for j in range(i + 1, self.n_atoms):
                pair = sy._get_pair(st.atoms[i].element, st.atoms[j].element)
                rep = sy.params[pair].get("repulsive") if pair in sy.params else None
                if rep is None or not callable(rep):
                    continue
                for im in range(st.max_image):
                    if st.bond_mat[im, i, j]:
                        e += float(rep(st.dist_mat[im, i, j]))
        return e

    def get_forces(self, *a, **k):
        raise NotImplementedError("the reference's Forces.get_forces raises TypeError as shipped (SURVEY.md); not reproduced")
