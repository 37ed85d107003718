"""numpy <-> PLY for the fused point cloud (reference: lib/ply_np_converter.py:38-92, which goes through a
vendored `plyfile`).  Writes/reads the subset the pipeline uses: one `vertex` element with x,y,z (float32 or
float64) and optional red,green,blue (uint8), binary little-endian or ASCII, with comment lines."""
import numpy as np


def np2ply(vertex, out_ply, color=None, comments=None, text=False, use_double=False):
    ftype, fname = ('<f8', 'double') if use_double else ('<f4', 'float')
    vertex = np.asarray(vertex)
    n = vertex.shape[0]
    fields = [('x', ftype), ('y', ftype), ('z', ftype)]
    if color is not None:
        fields += [('red', 'u1'), ('green', 'u1'), ('blue', 'u1')]
    data = np.empty(n, dtype=fields)
    data['x'], data['y'], data['z'] = vertex[:, 0], vertex[:, 1], vertex[:, 2]
    if color is not None:
        color = np.asarray(color)
        data['red'], data['green'], data['blue'] = color[:, 0], color[:, 1], color[:, 2]
    header = ['ply', 'format {} 1.0'.format('ascii' if text else 'binary_little_endian')]
    for c in (comments or []):
        header.append('comment {}'.format(c))
    header.append('element vertex {}'.format(n))
    header += ['property {} {}'.format(fname, a) for a in ('x', 'y', 'z')]
    if color is not None:
        header += ['property uchar {}'.format(a) for a in ('red', 'green', 'blue')]
    header.append('end_header')
    with open(out_ply, 'wb') as fp:
        fp.write(('\n'.join(header) + '\n').encode('ascii'))
        if text:
            for row in data:
                vals = ['%.4f' % row['x'], '%.4f' % row['y'], '%.4f' % row['z']]
                if color is not None:
                    vals += ['%i' % row['red'], '%i' % row['green'], '%i' % row['blue']]
                fp.write((' '.join(vals) + '\n').encode('ascii'))
        else:
            fp.write(data.tobytes())


def ply2np(in_ply):
    with open(in_ply, 'rb') as fp:
        fmt, n, props, comments = None, 0, [], []
        while True:
            line = fp.readline().decode('ascii').strip()
            if line.startswith('format'):
                fmt = line.split()[1]
            elif line.startswith('comment'):
                comments.append(line[len('comment '):])
            elif line.startswith('element vertex'):
                n = int(line.split()[2])
            elif line.startswith('property'):
                _, typ, name = line.split()
                props.append((name, {'float': 'f4', 'double': 'f8', 'uchar': 'u1', 'float32': 'f4', 'float64': 'f8',
                                     'uint8': 'u1'}[typ]))
            elif line == 'end_header':
                break
        if fmt == 'ascii':
            raw = np.loadtxt(fp, ndmin=2)
            cols = {name: raw[:, i] for i, (name, _) in enumerate(props)}
        else:
            bo = '<' if fmt == 'binary_little_endian' else '>'
            arr = np.frombuffer(fp.read(), dtype=[(nm, bo + t if t != 'u1' else t) for nm, t in props], count=n)
            cols = {name: arr[name] for name, _ in props}
    data = np.stack([cols['x'], cols['y'], cols['z']], axis=1) if 'x' in cols else None
    color = np.stack([cols['red'], cols['green'], cols['blue']], axis=1).astype(np.uint8) if 'red' in cols else None
    return data, color, (comments if comments else None)
