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
