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


_PLY_TYPES = {'char': 'i1', 'int8': 'i1', 'uchar': 'u1', 'uint8': 'u1', 'short': 'i2', 'int16': 'i2', 'ushort': 'u2',
              'uint16': 'u2', 'int': 'i4', 'int32': 'i4', 'uint': 'u4', 'uint32': 'u4', 'float': 'f4', 'float32': 'f4',
              'double': 'f8', 'float64': 'f8'}


def ply2np(in_ply):
    """(N,3) xyz, (N,3) uint8 colour or None, comments or None -- the return of lib/ply_np_converter.py:71-92.  Reads the
    `vertex` element (any scalar properties, e.g. COLMAP's x y z nx ny nz red green blue) of an ASCII or binary PLY;
    the vertex element must come first (it does in every writer of this pipeline); later elements (faces) are ignored."""
    with open(in_ply, 'rb') as fp:
        fmt, n, props, comments, element = None, 0, [], [], None
        if fp.readline().strip() != b'ply':
            raise ValueError('not a PLY file: {}'.format(in_ply))
        while True:
            raw = fp.readline()
            if not raw:
                raise ValueError('PLY header has no end_header: {}'.format(in_ply))
            line = raw.decode('ascii', 'replace').strip()
            if line.startswith('format'):
                fmt = line.split()[1]
            elif line.startswith('comment'):
                comments.append(line[len('comment '):])
            elif line.startswith('element'):
                element = line.split()[1]
                if element == 'vertex':
                    if props:
                        raise ValueError('two vertex elements')
                    n = int(line.split()[2])
                elif not props and int(line.split()[2]) > 0:
                    raise ValueError('PLY element {} precedes the vertices: {}'.format(element, in_ply))
            elif line.startswith('property') and element == 'vertex':
                tok = line.split()
                if tok[1] == 'list':
                    raise ValueError('list property in the vertex element: {}'.format(in_ply))
                props.append((tok[2], _PLY_TYPES[tok[1]]))
            elif line == 'end_header':
                break
        if fmt == 'ascii':
            rows = [fp.readline().split() for _ in range(n)]
            raw = np.array(rows, dtype=np.float64).reshape(n, len(props))
            cols = {name: raw[:, i] for i, (name, _) in enumerate(props)}
        else:
            bo = '<' if fmt == 'binary_little_endian' else '>'
            dt = np.dtype([(nm, t if t[1] == '1' else bo + t) for nm, t in props])
            arr = np.frombuffer(fp.read(n * dt.itemsize), dtype=dt, count=n)
            cols = {name: arr[name] for name, _ in props}
    data = np.stack([cols['x'], cols['y'], cols['z']], axis=1) if 'x' in cols else None
    color = np.stack([cols['red'], cols['green'], cols['blue']], axis=1).astype(np.uint8) if 'red' in cols else None
    return data, color, (comments if comments else None)
