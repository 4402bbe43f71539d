import sys, numpy as np, torch, scipy.sparse as sp
sys.path.insert(0, '.')
from oracle import nksr_oracle as O
from tests import clouds
import nksr_b200
cuda = torch.device('cuda:0')
xyz, nrm = clouds.shapenet_like(3000)
W, L, C = 0.02, 4, 4
svh = nksr_b200.SparseFeatureHierarchy(W, L, cuda).build_point_splatting(torch.from_numpy(xyz).to(cuda))
osvh = O.OracleSVH(W, L).build_point_splatting(xyz)
rng = np.random.default_rng(11)
feats = [(0.5 + 0.2 * rng.normal(size=(osvh.n(l), C))).astype(np.float32) for l in range(L)]
offs = osvh.offsets()
lev = np.zeros(offs[-1], int)
for l in range(L): lev[offs[l]:offs[l+1]] = l
t = lambda a: torch.from_numpy(a).to(cuda)
nxyz_all = np.concatenate([osvh.centers(0), osvh.centers(1)])
rng2 = np.random.default_rng(3)
nval_all = rng2.normal(size=nxyz_all.shape).astype(np.float32)
def run(name, pos, nx, nv, pw, nw, rw, approx=False):
    field = nksr_b200.KernelField(svh, None, [t(f) for f in feats], approx)
    field.solver_config.update(keep_system=True, max_iter=0)
    field.solve(t(pos), t(nx) if nx is not None else None, t(nv) if nv is not None else None, pw, nw, rw)
    s = field.system
    n = s.rowptr.numel() - 1
    A = sp.csr_matrix((s.val.cpu().numpy().astype(np.float64), s.col.cpu().numpy(), s.rowptr.cpu().numpy()), shape=(n, n))
    Ar, br, _ = O.build_system(osvh, feats, pos, nx if nx is not None else np.zeros((0,3),np.float32), nv if nv is not None else np.zeros((0,3)), pw, nw, rw, approx)
    D = (A - Ar).tocoo()
    print(name, 'scale', abs(Ar).max(), 'maxdiff', abs(D.data).max() if D.nnz else 0, 'rhs diff', np.abs(s.rhs.cpu().numpy()-br).max(), 'rhs scale', np.abs(br).max())
    for a in range(L):
        for b in range(L):
            m = (lev[D.row] == a) & (lev[D.col] == b)
            if m.any():
                mr = (lev[Ar.tocoo().row]==a)&(lev[Ar.tocoo().col]==b)
                print('   block', a, b, 'maxdiff %.3e' % abs(D.data[m]).max(), 'ref max %.3e' % (abs(Ar.tocoo().data[mr]).max() if mr.any() else 0), 'n bad', int((abs(D.data[m])>1e-3*abs(Ar).max()).sum()))
run('reg only', xyz[:0], None, None, 1.0, 0.0, 1.0)
run('pos only', xyz, None, None, 3.3, 0.0, 0.0)
run('nrm lvl0 only', xyz[:0], osvh.centers(0), nval_all[:osvh.n(0)], 1.0, 0.05, 0.0)
run('nrm lvl0 approx', xyz[:0], osvh.centers(0), nval_all[:osvh.n(0)], 1.0, 0.05, 0.0, True)
run('nrm lvl1 only', xyz[:0], osvh.centers(1), nval_all[osvh.n(0):], 1.0, 0.05, 0.0)
run('nrm offcentre', xyz[:0], (osvh.centers(0)+0.003).astype(np.float32), nval_all[:osvh.n(0)], 1.0, 0.05, 0.0)
