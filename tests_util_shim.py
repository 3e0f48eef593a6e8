"""bench.py helper: the oracle's RPC model from a 90-double table (same as tests/util.rpc_from_array)."""
from oracle import rpc_oracle


def rpc_from_array(a):
    r = rpc_oracle.RPCModel()
    (r.row_offset, r.col_offset, r.lat_offset, r.lon_offset, r.alt_offset,
     r.row_scale, r.col_scale, r.lat_scale, r.lon_scale, r.alt_scale) = [float(v) for v in a[:10]]
    r.row_num, r.row_den, r.col_num, r.col_den = [list(a[10 + 20 * i: 30 + 20 * i]) for i in range(4)]
    return r
