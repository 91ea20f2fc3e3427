"""The index rules of ``pad_kernel`` and ``ghost_fill_kernel`` (amx_api.cu), transcribed to Python loops
and checked against the oracle's ``pad`` / ``boundary_handler`` / ``prepare_state`` for every boundary
kind (open, reflective incl. the 1-D variant and the 3-vector flip rule Q12, periodic, mixed sides, the
MHD jet inflow).  The kernels themselves are compared with the oracle on the GPU
(``test_z_full_size_properties.py::test_layout_seams_vs_oracle``); this is the same comparison for the
addressing logic alone, runnable without a device: every ghost cell has one interior source cell (the
composition of the per-axis maps), one sign flip per reflective axis it lies beyond, and the jet
constants are written last."""
import itertools

import numpy as np
import pytest

from astronomix_b200 import option_classes as oc
from astronomix_b200 import problems as pr
from oracle import fv_oracle as fo
from test_z_full_size_properties import _boundaries

OPEN,REFL,PER,JET=0,1,2,3
def geom(config, n, dim):
    bc=oc.boundary_codes(config)
    g={'nd':dim,'n':[n]*dim,'bc':bc,'jet_lo':0,'jet_hi':0}
    if dim==2 and bc[1][0]==JET:
        h=int(0.025/config.grid_spacing); g['jet_lo']=n//2-h-2; g['jet_hi']=n//2+h-2
    return g
def jet_cell(g,ix,iy): return g['jet_hi']>g['jet_lo'] and ix<0 and g['jet_lo']<=iy<g['jet_hi']
def jet_gas(k): return [5/3,800.0,0.0,1.0][k]
def jet_vec(t,k): return 0.0 if k else (800.0 if t==1 else 200**0.5)
def ghost_fill(g, a, rows, type_handled, ng=2):
    # transcription of ghost_fill_kernel (in place; sources are interior cells)
    nd=g['nd']; out=a.copy()
    for idx in itertools.product(*[range(g['n'][k]+2*ng) for k in range(nd)]):
        flipmask=0; ghost=False; src=[]
        for ax in range(nd):
            ia=idx[ax]-ng; ig=ia
            if ia<0:
                ghost=True; b=g['bc'][ax][0]
                ig = ia+g['n'][ax] if b==PER else ((0 if nd==1 else -1-ia) if b==REFL else (0 if b in (OPEN,JET) else ia))
                if b==REFL: flipmask|=1<<(ax+1)
            elif ia>=g['n'][ax]:
                ghost=True; b=g['bc'][ax][1]
                ig = ia-g['n'][ax] if b==PER else ((g['n'][ax]-1 if nd==1 else 2*g['n'][ax]-1-ia) if b==REFL else (g['n'][ax]-1 if b==OPEN else ia))
                if b==REFL: flipmask|=1<<(ax+1)
            src.append(ig+ng)
        jet = nd==2 and jet_cell(g, idx[0]-ng, idx[1]-ng)
        if not ghost or (tuple(src)==idx and not jet): continue
        for k in range(rows):
            v=a[(k,)+tuple(src)]
            if flipmask&(1<<k): v=-v
            if jet: v = jet_gas(k) if type_handled==0 else jet_vec(type_handled,k)
            out[(k,)+idx]=v
    return out
def pad_k(g, s, rows, gas_rows, ng=2):
    nd=g['nd']; out=np.zeros((rows,)+tuple(n+2*ng for n in g['n']),dtype=s.dtype)
    for idx in itertools.product(*[range(g['n'][k]+2*ng) for k in range(nd)]):
        flipmask=0; sg=[]; se=[]
        for a in range(nd):
            ia=idx[a]-ng; n=g['n'][a]
            se.append(0 if ia<0 else (n-1 if ia>=n else ia)); ig=ia
            if ia<0:
                b=g['bc'][a][0]; ig = ia+n if b==PER else ((0 if nd==1 else -1-ia) if b==REFL else 0)
                if b==REFL: flipmask|=1<<(a+1)
            elif ia>=n:
                b=g['bc'][a][1]; ig = ia-n if b==PER else ((n-1 if nd==1 else 2*n-1-ia) if b==REFL else n-1)
                if b==REFL: flipmask|=1<<(a+1)
            sg.append(ig)
        for k in range(rows):
            if k<gas_rows:
                v=s[(k,)+tuple(sg)]
                if flipmask&(1<<k): v=-v
                if nd==2 and jet_cell(g,idx[0]-ng,idx[1]-ng): v=jet_gas(k)
            else: v=s[(k,)+tuple(se)]
            out[(k,)+idx]=v
    return out

CASES = [(1, 33, False, "open"), (1, 33, False, "reflective"), (1, 32, False, "periodic"), (2, 20, False, "mixed"),
         (2, 20, True, "reflective"), (2, 48, True, "jet"), (2, 48, False, "jet"), (3, 10, True, "mixed"),
         (3, 10, True, "reflective"), (3, 10, False, "periodic"), (3, 10, True, "open")]


@pytest.mark.parametrize("dim,n,mhd,kind", CASES)
def test_kernel_index_rules_reproduce_the_oracle(dim, n, mhd, kind):
    s = pr.smooth_random(dim, n, mhd=mhd, seed=3, dtype=np.float64)
    config, params, rv = pr.make(dim, n, mhd=mhd, boundaries=_boundaries(dim, kind), dtype=np.float64)
    o = fo.Opts.from_config(config, params)
    ng_ = dim + 2
    g = geom(config, n, dim)
    all_open = dict(g, bc=[(OPEN, OPEN)] * dim, jet_lo=0, jet_hi=0)  # what stepping._pad asks amx_pad for
    padded = pad_k(all_open, s, s.shape[0], ng_)
    assert np.array_equal(padded, fo.pad(s, o))
    assert np.array_equal(ghost_fill(g, padded[:ng_], ng_, 0), fo.boundary_handler(padded[:ng_], o, fo.GAS_STATE))
    assert np.array_equal(pad_k(g, s, s.shape[0], ng_)[:ng_], fo.prepare_state(s, o)[:ng_])  # amx_pad / shell start
    if mhd:
        vec = np.random.default_rng(4).standard_normal((3,) + padded.shape[1:])
        for kind_, code in ((fo.VELOCITY_ONLY, 1), (fo.MAGNETIC_FIELD_ONLY, 2)):
            assert np.array_equal(ghost_fill(g, vec, 3, code), fo.boundary_handler(vec, o, kind_))
