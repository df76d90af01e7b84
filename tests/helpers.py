"""Shared helpers for the tests: golden-case unpacking on bare arrays."""
import json

import numpy as np


def split_field(shape, f):
    """Views (fx, fy, fz) of a 1-D field for cell shape (nx, ny, nz)."""
    nx, ny, nz = shape
    shp = ((nx, ny + 1, nz + 1), (nx + 1, ny, nz + 1), (nx + 1, ny + 1, nz))
    out, i0 = [], 0
    for s in shp:
        n = int(np.prod(s))
        out.append(f[i0:i0 + n].reshape(s, order='F'))
        i0 += n
    return out


def kernel_case(gk, k):
    """Inputs of kernel golden case k as a dict."""
    p = f"k{k}_"
    keys = ('hx', 'hy', 'hz', 'eta_x', 'eta_y', 'eta_z', 'zeta', 's', 'e', 'r')
    d = {n: gk[p + n] for n in keys}
    d['shape'] = (d['hx'].size, d['hy'].size, d['hz'].size)
    for n in ('eta_x', 'eta_y', 'eta_z', 'zeta'):
        d[n] = np.asfortranarray(d[n])
    d['prefix'] = p
    return d


def solve_case(gs, prefix):
    """Inputs/outputs of one golden solve."""
    d = {}
    for n in ('hx', 'hy', 'hz', 'origin', 'sfield', 'efield', 'error_at_cycle'):
        d[n] = gs[prefix + n]
    d['model'] = {n: gs[prefix + n] for n in
                  ('property_x', 'property_y', 'property_z', 'mu_r', 'epsilon_r')
                  if prefix + n in gs.files}
    d['frequency'] = float(gs[prefix + 'frequency'])
    d['kwargs'] = json.loads(str(gs[prefix + 'kwargs']))
    d['it_mg'] = int(gs[prefix + 'it_mg'])
    d['it_ssl'] = int(gs[prefix + 'it_ssl'])
    d['exit_message'] = str(gs[prefix + 'exit_message'])
    d['abs_error'] = float(gs[prefix + 'abs_error'])
    d['ref_error'] = float(gs[prefix + 'ref_error'])
    d['stdout'] = str(gs[prefix + 'stdout'])
    d['source'] = gs[prefix + 'source'] if prefix + 'source' in gs.files else None
    return d


def split_faces(shape, f):
    """Views (hx, hy, hz) of a 1-D magnetic (face) field for cell shape (nx, ny, nz)."""
    nx, ny, nz = shape
    shp = ((nx + 1, ny, nz), (nx, ny + 1, nz), (nx, ny, nz + 1))
    out, i0 = [], 0
    for s in shp:
        n = int(np.prod(s))
        out.append(f[i0:i0 + n].reshape(s, order='F'))
        i0 += n
    return out


def hfield_case(gh, k):
    """Inputs / outputs of magnetic-field golden case k."""
    p = f"h{k}_"
    d = {n: gh[p + n] for n in ('hx', 'hy', 'hz', 'origin', 'property_x', 'property_y',
                                'property_z', 'mu_r', 'e', 'h', 'zeta_k', 'h_k')}
    d['frequency'] = float(gh[p + 'frequency'])
    d['shape'] = (d['hx'].size, d['hy'].size, d['hz'].size)
    return d


def maps_grid(gm, prefix, suffix=''):
    """(h list, origin) of a grid stored in maps.npz as concatenated widths + cell counts."""
    key_h = {'': 'h', '_in': 'h_in', '_out': 'h_out', '2': 'h2'}[suffix]
    key_n = {'': 'n', '_in': 'n_in', '_out': 'n_out', '2': 'n2'}[suffix]
    key_o = {'': 'origin', '_in': 'o_in', '_out': 'o_out', '2': 'origin2'}[suffix]
    h, n = gm[prefix + key_h], [int(v) for v in gm[prefix + key_n]]
    hs = [h[:n[0]], h[n[0]:n[0] + n[1]], h[n[0] + n[1]:]]
    origin = gm[prefix + key_o] if prefix + key_o in gm.files else np.zeros(3)
    return hs, origin
