"""Pure-Python stand-in for dm-tree (TEST INFRASTRUCTURE ONLY).

The reference's vima/utils.py:4 does `import tree` (dm-tree, a C++ extension that is not in this
image).  Only the five functions below are used (vima/utils.py:44,64,311,340,502,525,544,693,739,
809,813,831,871,875).  Mappings are traversed in sorted-key order and rebuilt with their own type;
sequences are rebuilt with their own type -- the same conventions dm-tree follows.
"""
import collections.abc as _abc


def _is_mapping(x):
    return isinstance(x, _abc.Mapping)


def _is_seq(x):
    return isinstance(x, (list, tuple))


def _is_nest(x):
    return _is_mapping(x) or _is_seq(x)


def _rebuild(proto, children):
    if _is_mapping(proto):
        keys = sorted(proto.keys())
        out = dict(zip(keys, children))
        try:
            return type(proto)(out)
        except Exception:
            return out
    if isinstance(proto, tuple) and hasattr(proto, "_fields"):
        return type(proto)(*children)
    return type(proto)(children)


def _children(x):
    if _is_mapping(x):
        return [(k, x[k]) for k in sorted(x.keys())]
    return list(enumerate(x))


def flatten(s):
    if not _is_nest(s):
        return [s]
    out = []
    for _, c in _children(s):
        out.extend(flatten(c))
    return out


def _unflatten(proto, it):
    if not _is_nest(proto):
        return next(it)
    return _rebuild(proto, [_unflatten(c, it) for _, c in _children(proto)])


def unflatten_as(structure, flat_sequence):
    return _unflatten(structure, iter(list(flat_sequence)))


def map_structure_with_path(func, *structures, **kwargs):
    def rec(path, *xs):
        x0 = xs[0]
        if not _is_nest(x0):
            return func(path, *xs)
        kids = []
        for k, _ in _children(x0):
            kids.append(rec(path + (k,), *[x[k] for x in xs]))
        return _rebuild(x0, kids)

    return rec((), *structures)


def map_structure(func, *structures, **kwargs):
    return map_structure_with_path(lambda _p, *xs: func(*xs), *structures)


def traverse(fn, structure, top_down=True):
    def rec(x):
        if top_down:
            r = fn(x)
            if r is not None:
                return r
        if _is_nest(x):
            x = _rebuild(x, [rec(c) for _, c in _children(x)])
        if not top_down:
            r = fn(x)
            if r is not None:
                return r
        return x

    return rec(structure)
