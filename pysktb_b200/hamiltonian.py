"""`Hamiltonian`: the reference's k-point solver surface (pysktb/hamiltonian.py:59-317) over libsktb.so.

Same constructor, attributes, method names, argument meaning, return layouts and error text as the
reference; every numerical call goes to the sm_100a kernels through the C ABI (include/sktb.h):

    get_ham            -> sktb_hk           (K1, hamiltonian.py:98-106 + calc_g :275-317 / jit_modules.py:11-23)
    _sol_ham           -> sktb_heev         (K2, hamiltonian.py:108-126, clean_eig :223-231)
    solve_kpath/solve_k-> sktb_solve_host   (K1+K2 fused, chunked + stream-overlapped; hamiltonian.py:128-221)
    solve_kmesh (new)  -> sktb_solve_mesh   (K0+K1+K2, results stay on the device - for 512^3-class grids)
    get_kpts           -> System.get_kpts   (host, bit-exact; system.py:51-103)

Model construction (k-independent) stays on the host in numpy except the Slater-Koster tables, which
the plan evaluates on the device (K4).  There is no CPU fallback: without libsktb.so + a CUDA device the
constructor raises.

Deliberate deviations from the reference (documented in DESIGN.md):
  * `parallel=1, soc=False` works (the reference crashes with a broadcast ValueError);
    `parallel` only selects between "solve" (0/1/True/False), the reference's Exception (2) and None (other).
  * the stray `print(2)` of hamiltonian.py:197 is not reproduced.
"""
import os

import numpy as np

from . import _lib
from .atom import Atom
from .system import System

_NOT_HERMITIAN = "\n\nHamiltonian matrix is not hermitian?!"
_PIN_THRESHOLD = 1 << 20


def _torch():
    import torch

    if not torch.cuda.is_available():
        raise _lib.SktbError("no CUDA device visible: pysktb_b200 has no CPU fallback")
    return torch


def host_buffer(shape, dtype):
    """numpy array for results; page-locked (through torch) when large so D2H copies overlap compute."""
    dtype = np.dtype(dtype)
    nbytes = int(np.prod(shape)) * dtype.itemsize
    if nbytes >= _PIN_THRESHOLD:
        torch = _torch()
        tdt = {np.dtype(np.float64): torch.float64, np.dtype(np.complex128): torch.complex128,
               np.dtype(np.int32): torch.int32}[dtype]
        try:
            return torch.empty(tuple(int(s) for s in shape), dtype=tdt, pin_memory=True).numpy()
        except RuntimeError:
            pass
    return np.empty(shape, dtype=dtype)


class Hamiltonian(object):
    """Object to represent hamiltonian of a system (reference: hamiltonian.py:59)."""

    E_PREFIX = "e_"

    def __init__(self, structure, inter, numba=1, device=None):
        self.structure = structure
        self.inter = inter
        t = {a.element: a.orbitals for a in self.structure.atoms}          # hamiltonian.py:68
        self.system = System(self.structure, t, self.inter)
        self.n_orbitals = len(self.system.all_orbitals)
        self.soc_mat = self.system.get_soc_mat()
        self.dist_mat_vec = self.system.structure.dist_mat_vec
        self.bond_mat = self.system.structure.bond_mat
        self.numba = numba                                                  # accepted, unused: the CUDA path is the only path
        self.orbital_order = self._orbital_order()
        self._H_wo_g = None
        self._plan = None
        self._device = device
        self._build_plan()

    # -- plan ---------------------------------------------------------------------------------------
    def _build_plan(self):
        torch = _torch()
        if self._device is None:
            self._device = int(os.environ.get("LOCAL_RANK", torch.cuda.current_device())) % torch.cuda.device_count()
        st, sy = self.structure, self.system
        names = _lib.v_names()
        col = {n: i for i, n in enumerate(names)}
        bl = st.bond_list()
        nb = len(bl)
        V = np.zeros((nb, len(names)))
        vec = np.zeros((nb, 3))
        lmn = np.zeros((nb, 3))
        const_cache = {}
        for b, (im, i, j) in enumerate(bl):
            pair = (st.atoms[i].element, st.atoms[j].element)
            if pair in const_cache:
                hop = const_cache[pair]
            else:
                hop = sy.get_hop_params(i, j, im)
                from .scaling import ScalingLaw
                raw = sy.params[sy._get_pair(*pair)]
                if not any(isinstance(v, ScalingLaw) for v in raw.values()):
                    const_cache[pair] = hop
            for key, val in hop.items():
                if key not in col:   # the reference passes **params to get_hop_int -> TypeError on unknown names
                    raise TypeError(f"get_hop_int() got an unexpected keyword argument '{key}'")
                V[b, col[key]] = val
            d = st.dist_mat_vec[im, i, j, :]
            vec[b] = d
            lmn[b] = d / np.linalg.norm(d)                                  # structure.py:168-174
        self._bonds = bl
        starts = np.cumsum([0] + [len(a.orbitals) for a in st.atoms])
        kinds = [Atom.ORBITALS_ALL.index(o) for (_, _, _, o) in sy.all_iter]
        onsite = np.concatenate([np.diag(sy.get_onsite_term(a)).astype(float) for a in range(len(st.atoms))])
        r, c = np.nonzero(self.soc_mat)
        sv = self.soc_mat[r, c]
        keep = dict(
            start=_lib.as_i32(starts), kind=_lib.as_i32(kinds), onsite=_lib.as_f64(onsite),
            img=_lib.as_i32(bl[:, 0] if nb else []), bi=_lib.as_i32(bl[:, 1] if nb else []), bj=_lib.as_i32(bl[:, 2] if nb else []),
            vec=_lib.as_f64(vec), lmn=_lib.as_f64(lmn), V=_lib.as_f64(V),
            rec=_lib.as_f64(st.lattice.get_rec_lattice()),
            sr=_lib.as_i32(r), sc=_lib.as_i32(c), sv=_lib.as_f64(np.stack([sv.real, sv.imag], axis=1) if len(sv) else np.zeros((0, 2))),
        )
        d = _lib.ModelDesc(
            n_atoms=len(st.atoms), n_orb=self.n_orbitals,
            atom_orb_start=_lib.ptr_i32(keep["start"]), orb_kind=_lib.ptr_i32(keep["kind"]), onsite=_lib.ptr_f64(keep["onsite"]),
            n_bonds=nb, bond_image=_lib.ptr_i32(keep["img"]), bond_i=_lib.ptr_i32(keep["bi"]), bond_j=_lib.ptr_i32(keep["bj"]),
            bond_vec=_lib.ptr_f64(keep["vec"]), bond_lmn=_lib.ptr_f64(keep["lmn"]), bond_V=_lib.ptr_f64(keep["V"]),
            rec_lattice=_lib.ptr_f64(keep["rec"]),
            soc_nnz=len(r), soc_row=_lib.ptr_i32(keep["sr"]), soc_col=_lib.ptr_i32(keep["sc"]), soc_val=_lib.ptr_f64(keep["sv"]),
        )
        handle = _lib.C.c_void_p()
        _lib.check(_lib.lib().sktb_plan_create(_lib.C.byref(d), self._device, _lib.C.byref(handle)))
        self._plan = handle
        self._starts = starts
        self._onsite = onsite
        bound = _lib.C.c_double()
        _lib.check(_lib.lib().sktb_plan_asym_bound(self._plan, _lib.C.byref(bound)))
        self.asym_bound = bound.value
        defect = _lib.C.c_double(0.0)
        _lib.check(_lib.lib().sktb_plan_herm_defect(self._plan, 1, _lib.C.byref(defect)))
        self.herm_defect = defect.value            # 0 (to rounding) <=> H(k) is Hermitian for every k (greens.py picks its path by it)
        _lib.check(_lib.lib().sktb_plan_herm_defect(self._plan, 0, _lib.C.byref(defect)))
        self.herm_defect_nosoc = defect.value      # the same for get_ham(k, l_soc=False): hopping blocks only

    def __del__(self):
        try:
            if getattr(self, "_plan", None) is not None and _lib._lib is not None:
                _lib._lib.sktb_plan_destroy(self._plan)
                self._plan = None
        except Exception:
            pass

    def __getstate__(self):   # the reference pickles Hamiltonians to joblib workers; plans are per process
        s = dict(self.__dict__)
        s["_plan"] = None
        return s

    def __setstate__(self, s):
        self.__dict__.update(s)
        self._build_plan()

    @property
    def H_wo_g(self):
        """(max_image, n, n) complex hopping tensor of the reference (hamiltonian.py:71-74, 425-488),
        rebuilt on demand from the plan's device-computed Slater-Koster blocks."""
        if self._H_wo_g is None:
            n, mi = self.n_orbitals, int(self.structure.max_image)
            sizes = [(self._starts[i + 1] - self._starts[i]) * (self._starts[j + 1] - self._starts[j]) for _, i, j in self._bonds]
            flat = np.zeros(int(np.sum(sizes)) if len(sizes) else 0)
            _lib.check(_lib.lib().sktb_plan_hoppings(self._plan, flat.ctypes.data, flat.size))
            H = np.zeros((mi, n, n), dtype=complex)
            at = 0
            for (im, i, j), sz in zip(self._bonds, sizes):
                ri, rj = slice(self._starts[i], self._starts[i + 1]), slice(self._starts[j], self._starts[j + 1])
                H[im, ri, rj] = flat[at:at + sz].reshape(ri.stop - ri.start, rj.stop - rj.start)
                at += sz
            c = int(mi / 2)
            for a in range(len(self.structure.atoms)):
                blk = slice(self._starts[a], self._starts[a + 1])
                H[c, blk, blk] = np.diag(self._onsite[blk])
            self._H_wo_g = H
        return self._H_wo_g

    # -- small helpers ---------------------------------------------------------------------------
    @staticmethod
    def get_orb_ind(orbit):
        return Atom.ORBITALS_ALL.index(orbit)

    def _orbital_order(self):
        out, k = {}, 0
        for el, orb in self.system.all_orbitals:
            out[k] = f"{el}-{orb}-up"
            out[k + 1] = f"{el}-{orb}-down"
            k += 2
        return out

    def _dim(self, soc):
        return self.n_orbitals * 2 if soc else self.n_orbitals

    def _stream(self):
        return _torch().cuda.current_stream(self._device).cuda_stream

    def get_kpts(self, path, nk):
        return self.system.get_kpts(path, nk)

    def k_cart2red(self, k):
        red2cart = np.array([self.structure.get_lattice()[i][: len(k)] for i in range(len(k))]).transpose()
        return np.linalg.inv(red2cart) @ np.array(k)

    def k_red2cart(self, k):
        red2cart = np.array([self.structure.get_lattice()[i][: len(k)] for i in range(len(k))]).transpose()
        return red2cart @ np.array(k)

    # -- K1 ----------------------------------------------------------------------------------------
    def get_ham(self, kpt, l_soc=True):
        """H(k) as the reference returns it: (2n,2n) with SOC, (n,n) otherwise (hamiltonian.py:98-106)."""
        return self.get_ham_batch(np.asarray(kpt, dtype=float).reshape(1, 3), l_soc)[0]

    def get_ham_batch(self, k_list, l_soc=True, return_flags=False):
        torch = _torch()
        soc = bool(l_soc == True)  # noqa: E712 - the reference tests `l_soc == True`
        k = np.ascontiguousarray(np.asarray(k_list, dtype=float).reshape(-1, 3))
        N = self._dim(soc)
        with torch.cuda.device(self._device):
            kd = torch.from_numpy(k).cuda()
            H = torch.empty((len(k), N, N), dtype=torch.complex128, device="cuda")
            fl = torch.zeros(len(k), dtype=torch.int32, device="cuda")
            _lib.check(_lib.lib().sktb_hk(self._plan, kd.data_ptr(), len(k), int(soc), H.data_ptr(), fl.data_ptr(), self._stream()))
            out = H.cpu().numpy()
            flags = fl.cpu().numpy()
        return (out, flags) if return_flags else out

    def _gf_inverse_raw(self, H, energies, eta, want="G"):
        """Device inverse of caller-supplied matrices (sktb_gf_inverse; diagnostics / tests): H (nm, N, N) or (N, N), energies (nE,).
        want="G": inv((E + i eta) I - H), shape (nm, nE, N, N) - shared-memory kernel; want="trace": -(1/pi) Im Tr G, shape (nm, nE) -
        register-resident kernels."""
        torch = _torch()
        H = np.ascontiguousarray(np.asarray(H, dtype=complex))
        H = H.reshape((-1,) + H.shape[-2:])
        nm, n = H.shape[0], H.shape[1]
        E = np.ascontiguousarray(np.atleast_1d(np.asarray(energies, dtype=float)))
        with torch.cuda.device(self._device):
            Hd, Ed = torch.from_numpy(H).cuda(), torch.from_numpy(E).cuda()
            if want == "G":
                out = torch.empty((nm, len(E), n, n), dtype=torch.complex128, device="cuda")
                _lib.check(_lib.lib().sktb_gf_inverse(self._plan, Hd.data_ptr(), n, nm, Ed.data_ptr(), len(E), float(eta), out.data_ptr(), None, self._stream()))
            else:
                out = torch.empty((nm, len(E)), dtype=torch.float64, device="cuda")
                _lib.check(_lib.lib().sktb_gf_inverse(self._plan, Hd.data_ptr(), n, nm, Ed.data_ptr(), len(E), float(eta), None, out.data_ptr(), self._stream()))
            return out.cpu().numpy()

    def calc_g(self, kpt):
        """Phase tensor g[image, i, j] of the reference (hamiltonian.py:275-317).  Diagnostic only: the
        kernels never materialise it (one phase per bond, not per orbital pair); built here on the host
        from the bond list with the reference's arithmetic."""
        n, mi = self.n_orbitals, int(self.structure.max_image)
        kc = np.dot(kpt, self.structure.lattice.get_rec_lattice())
        g = np.zeros((mi, n, n), dtype=complex)
        for im, i, j in self._bonds:
            ph = np.exp(2.0 * np.pi * 1j * np.dot(kc, self.dist_mat_vec[im, i, j, :]))
            g[im, self._starts[i]:self._starts[i + 1], self._starts[j]:self._starts[j + 1]] = ph
        g[int(mi / 2), :, :] += np.eye(n, dtype=complex)
        return g

    # -- K2 ----------------------------------------------------------------------------------------
    def _sol_ham(self, ham, eig_vectors=False, spin=False):
        """Eigen-solve one matrix with the reference's gate, lower-triangle reading and ordering
        (hamiltonian.py:108-126)."""
        torch = _torch()
        ham = np.ascontiguousarray(np.asarray(ham, dtype=complex))
        N = ham.shape[0]
        with torch.cuda.device(self._device):
            Hd = torch.from_numpy(ham.reshape(1, N, N)).cuda()
            ev = torch.empty((N, 1), dtype=torch.float64, device="cuda")
            vec = torch.empty((N, 1, N), dtype=torch.complex128, device="cuda") if eig_vectors else None
            fl = torch.zeros(1, dtype=torch.int32, device="cuda")
            _lib.check(_lib.lib().sktb_heev(self._plan, Hd.data_ptr(), N, 1, ev.data_ptr(), vec.data_ptr() if eig_vectors else None,
                                            fl.data_ptr(), self._stream()))
            if int(fl.cpu()[0]):
                raise Exception(_NOT_HERMITIAN)
            vals = ev.cpu().numpy()[:, 0]
            if eig_vectors:
                return vals, vec.cpu().numpy()[:, 0, :]
        return vals

    # -- K1+K2 -------------------------------------------------------------------------------------
    def solve_kpath(self, k_list=None, eig_vectors=False, soc=True, parallel=1, out=None):
        """Eigenvalues [band, k] (and eigenvectors [band, k, component]) along a k list
        (hamiltonian.py:128-209).  `k_list` may also be a C-contiguous float64 (nk, 3) ndarray (used as is, no
        copy - page-locked arrays from `host_buffer` make the transfers asynchronous); `out` optionally supplies
        the (N, nk) float64 result array to reuse (extension for repeated large grids)."""
        if parallel == 2:
            raise Exception("\n\nparallel=2 not ready yet. please use parallel=1")
        if k_list is None or parallel not in (0, 1):
            return None
        soc = bool(soc == True)  # noqa: E712
        if isinstance(k_list, np.ndarray) and k_list.dtype == np.float64 and k_list.ndim == 2 and k_list.flags.c_contiguous:
            kh = k_list
        else:
            k = np.asarray([np.asarray(x, dtype=float) for x in k_list], dtype=float).reshape(-1, 3)
            kh = host_buffer(k.shape, np.float64)
            kh[...] = k
        nk, N = len(kh), self._dim(soc)
        if out is not None:
            assert out.shape == (N, nk) and out.dtype == np.float64 and out.flags.c_contiguous
        vals = out if out is not None else host_buffer((N, nk), np.float64)
        vecs = host_buffer((N, nk, N), np.complex128) if eig_vectors else None
        bad = _lib.C.c_int32(0)
        _lib.check(_lib.lib().sktb_solve_host(self._plan, kh.ctypes.data, nk, int(soc), vals.ctypes.data,
                                              vecs.ctypes.data if eig_vectors else None, _lib.C.byref(bad)))
        if bad.value:
            raise Exception(_NOT_HERMITIAN)
        self._check_convergence()
        return (vals, vecs) if eig_vectors else vals

    def solve_k(self, k_point=None, eig_vectors=False):
        if k_point is None:
            return None
        if not eig_vectors:
            return self.solve_kpath([k_point], eig_vectors=False)[:, 0]
        vals, vecs = self.solve_kpath([k_point], eig_vectors=True)
        return vals[:, 0], vecs[:, 0, :]

    def solve_kmesh(self, nk, soc=True, first=0, count=None, out=None, check=True):
        """Device-resident variant for large uniform meshes (GreensFunction._generate_kmesh order,
        greens.py:100-111): eigenvalues of mesh points [first, first+count) as a CUDA tensor (N, count).
        k-points are generated inside the kernel; nothing crosses PCIe.  (Extension: the reference has no
        equivalent because it cannot hold such grids.)"""
        torch = _torch()
        soc = bool(soc == True)  # noqa: E712
        nk3 = _lib.as_i32(list(nk))
        total = int(nk3[0]) * int(nk3[1]) * int(nk3[2])
        count = total - first if count is None else count
        N = self._dim(soc)
        with torch.cuda.device(self._device):
            if out is None:
                out = torch.empty((N, count), dtype=torch.float64, device="cuda")
            assert out.is_cuda and out.dtype == torch.float64 and out.shape[0] == N and out.stride(1) == 1
            gate = check and self.asym_bound > 0.5e-9
            fl = torch.zeros(count, dtype=torch.int32, device="cuda") if gate else None
            _lib.check(_lib.lib().sktb_solve_mesh(self._plan, _lib.ptr_i32(nk3), first, count, int(soc), out.data_ptr(),
                                                  out.stride(0), 0, fl.data_ptr() if gate else None, self._stream()))
            if gate and bool(fl.any().item()):
                raise Exception(_NOT_HERMITIAN)
        return out

    def get_dos(self, energy, eig=None, w=1e-2, nk=[20, 20, 20]):
        """Gaussian-broadened DOS (hamiltonian.py:233-253): eigenvalues on the endpoint-INCLUDED mesh
        `linspace(0, 1, nk[i])` (kx outer, kz inner) through `solve_kpath` (soc=True as in `solve_k`), or the
        caller's `eig`; sum_i exp(-(E - e_i)^2 / (2 w^2)) / (pi w sqrt(2)) on the device (sktb_gauss_dos)."""
        torch = _torch()
        energy = np.asarray(energy, dtype=float)
        with torch.cuda.device(self._device):
            if eig is None:
                kx, ky, kz = (np.linspace(0, 1, int(n)) for n in nk)
                mesh = np.stack(np.meshgrid(kx, ky, kz, indexing="ij"), axis=-1).reshape(-1, 3)
                kd = torch.from_numpy(np.ascontiguousarray(mesh)).cuda()
                N = self._dim(True)
                ev = torch.empty((N, len(mesh)), dtype=torch.float64, device="cuda")
                gate = self.asym_bound > 0.5e-9
                fl = torch.zeros(len(mesh), dtype=torch.int32, device="cuda") if gate else None
                _lib.check(_lib.lib().sktb_solve(self._plan, kd.data_ptr(), len(mesh), 1, ev.data_ptr(), None, len(mesh), 0,
                                                 fl.data_ptr() if gate else None, self._stream()))
                if gate and bool(fl.any().item()):
                    raise Exception(_NOT_HERMITIAN)
            else:
                ev = torch.from_numpy(np.ascontiguousarray(np.asarray(eig, dtype=float).reshape(-1))).cuda()
            Ed = torch.from_numpy(np.ascontiguousarray(energy.reshape(-1))).cuda()
            out = torch.empty(Ed.numel(), dtype=torch.float64, device="cuda")
            _lib.check(_lib.lib().sktb_gauss_dos(self._plan, ev.data_ptr(), ev.numel(), Ed.data_ptr(), Ed.numel(), float(w),
                                                 out.data_ptr(), self._stream()))
            res = out.cpu().numpy().reshape(energy.shape)
        return res

    def total_energy(self, filled_band=0, nk=10, dim=3, soc=True):
        """Mean of the `filled_band` lowest eigenvalues (default: half filling, N // 2) over the endpoint-included mesh
        linspace(0, 1, nk)^dim (hamiltonian.py:319-323 -> utils/energitics.py:26-63, which asks ARPACK for the lowest
        eigenvalues per k).  All eigenvalues come from the device solver; the partial sum is reduced on the device
        (sktb_band_energy).  'Caution for metallic systems: the Fermi level is not calculated' applies as in the reference."""
        torch = _torch()
        soc = bool(soc == True)  # noqa: E712
        x = np.linspace(0, 1, int(nk))
        if dim == 3:
            mesh = np.stack(np.meshgrid(x, x, x, indexing="ij"), axis=-1).reshape(-1, 3)
        elif dim == 2:
            mesh = np.concatenate([np.stack(np.meshgrid(x, x, indexing="ij"), axis=-1).reshape(-1, 2), np.zeros((int(nk) ** 2, 1))], axis=1)
        elif dim == 1:
            mesh = np.concatenate([x.reshape(-1, 1), np.zeros((int(nk), 2))], axis=1)
        else:
            raise ValueError("dim must be 1, 2 or 3")
        N = self._dim(soc)
        neig = int(filled_band) if filled_band else int(N / 2)
        neig = max(1, min(neig, N))
        with torch.cuda.device(self._device):
            kd = torch.from_numpy(np.ascontiguousarray(mesh)).cuda()
            ev = torch.empty((N, len(mesh)), dtype=torch.float64, device="cuda")
            gate = self.asym_bound > 0.5e-9
            fl = torch.zeros(len(mesh), dtype=torch.int32, device="cuda") if gate else None
            lib = _lib.lib()
            _lib.check(lib.sktb_solve(self._plan, kd.data_ptr(), len(mesh), int(soc), ev.data_ptr(), None, len(mesh), 0,
                                      fl.data_ptr() if gate else None, self._stream()))
            out = torch.zeros(1, dtype=torch.float64, device="cuda")
            _lib.check(lib.sktb_band_energy(self._plan, ev.data_ptr(), len(mesh), len(mesh), neig, out.data_ptr(), self._stream()))
            if gate and bool(fl.any().item()):    # the reference goes through ARPACK here, not through _sol_ham's gate
                import warnings
                warnings.warn("H(k) is not Hermitian for some k: total_energy follows the lower-triangle spectrum", RuntimeWarning)
            total = float(out.cpu()[0])
        return total / (len(mesh) * neig)

    @staticmethod
    def kproj_weights(vecs, index):
        """Projection weights of `plot_kproj` (hamiltonian.py:347-351): ||vecs[band, k, index]||_2, shape (band, k)."""
        v = np.asarray(vecs)[:, :, list(index)]
        return np.sqrt((np.abs(v) ** 2).sum(axis=-1))

    def clean_eig(self, eigen_val, eig=None):
        eigen_val = np.array(np.asarray(eigen_val).real, dtype=float)
        order = eigen_val.argsort()
        if eig is not None:
            return eigen_val[order], eig[order]
        return eigen_val[order]

    def ql_failures(self):
        """Number of matrices whose tridiagonal QL hit its iteration limit since the plan was created (synchronises).
        numpy raises LinAlgError in that situation (hamiltonian.py:119 -> zheevd INFO > 0); see `_check_convergence`."""
        n = _lib.C.c_int32(0)
        _lib.check(_lib.lib().sktb_ql_failures(self._plan, _lib.C.byref(n)))
        return int(n.value)

    def _check_convergence(self):
        n = self.ql_failures()
        if n > getattr(self, "_ql_seen", 0):
            self._ql_seen = n
            raise np.linalg.LinAlgError("Eigenvalues did not converge")

    def last_kernel_info(self):
        return _lib.lib().sktb_last_kernel_info().decode()
