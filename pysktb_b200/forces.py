"""Total energy of a tight-binding model (reference: pysktb/forces.py:367-425).

`get_total_energy`: the band energy runs on the device (mesh eigenvalues through the C ABI, reduced by
`sktb_band_energy`), the repulsive pair energy is summed on the host exactly as the reference does.

`get_forces`: the reference's implementation raises TypeError as shipped, so its
semantics are taken from its docstrings (forces.py:3-4, 43-58): F[a] = -dE_tot/dR_a in eV/Angstrom, shape
(n_atoms, 3), E_tot = E_band + E_rep of `get_total_energy`.  Here the derivative is a central difference of that
total energy over Cartesian displacements of one atom (6 n_atoms mesh solves, each one k-sharded device call), which
includes every dependence of H(k) on the positions: scaling laws, direction cosines and the Bloch phases.
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

    def get_forces(self, n_electrons, nk=[10, 10, 10], temperature=0.0, mu=None, soc=True, parallel=True, step=1e-4):
        """-> f64 (n_atoms, 3), eV/Angstrom (forces.py:43-72 signature; `parallel` is accepted and ignored, the mesh
        solve always runs on the device).  Zero electronic temperature only; `step` is the displacement in Angstrom
        of the central difference (error O(step^2))."""
        if temperature != 0.0 or mu is not None:
            raise NotImplementedError("get_forces: only temperature=0 with the filling fixed by n_electrons")
        from .atom import Atom
        from .structure import Structure
        from .hamiltonian import Hamiltonian
        st, ham = self.structure, self.ham
        lat_t = np.asarray(st.lattice.get_matrix(), dtype=float).T          # cart = lat.T @ frac (structure.py:104-143)
        forces = np.zeros((self.n_atoms, 3))
        for a in range(self.n_atoms):
            for d in range(3):
                e = []
                for sgn in (1.0, -1.0):
                    dcart = np.zeros(3)
                    dcart[d] = sgn * step
                    dfrac = np.linalg.solve(lat_t, dcart)
                    atoms = [Atom(x.element, np.asarray(x.pos, dtype=float) + (dfrac if i == a else 0.0), x.orbitals)
                             for i, x in enumerate(st.atoms)]
                    st2 = Structure(st.lattice, atoms, periodicity=st.periodicity, name=st.name, bond_cut=st.bond_cut,
                                    numba=st.numba)
                    ham2 = Hamiltonian(st2, ham.inter, numba=ham.numba, device=ham._device)
                    e.append(Forces(ham2).get_total_energy(n_electrons, nk=nk, soc=soc)[0])
                    del ham2
                forces[a, d] = -(e[0] - e[1]) / (2.0 * step)
        return forces
