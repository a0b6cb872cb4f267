"""Python restatement of the order-free ICAO filter evaluation that the CUDA
path implements in events_finalize_kernel / resolve_kernel; used by CPU tests to check the
algorithm (against the sequential oracle) and the sharding protocol (on gloo)."""
K_NONE, K_PS, K_11Z, K_11I, K_17, K_18, K_PL = range(7)
NT = 1 << 25


def local_events(records_per_buffer, ordinals):
    """first-add candidates {key: (buffer ordinal, j, t)} of a set of buffers."""
    first = {}
    for recs, ob in zip(records_per_buffer, ordinals):
        for j, w in recs:
            for t, wd in enumerate(w):
                kind, key = wd >> 29, wd & 0xFFFFFF
                if kind in (K_11Z, K_17, K_18) and key != 0:
                    k = key | (NT if kind == K_18 else 0)
                    o = (ob, j, t)
                    if k not in first or o < first[k]:
                        first[k] = o
    return first


def merge_events(dicts):
    out = {}
    for d in dicts:
        for k, o in d.items():
            if k not in out or o < out[k]:
                out[k] = o
    return out


def finalize(first, preloaded=frozenset(), capacity=4096):
    """Admitted keys {key: first ordinal} after the DF18 rule and the 4096-entry capacity."""
    first = dict(first)
    for k in [k for k in first if k >> 25]:
        plain = k & 0xFFFFFF
        if plain in preloaded or (plain in first and first[plain] < first[k]):
            del first[k]          # icao_filter_test(addr) was already true: DF18 adds nothing
    new = sorted((o, k) for k, o in first.items() if k not in preloaded)
    room = max(capacity - len(preloaded), 0)
    return {k: o for o, k in new[:room]}


def resolve(records_per_buffer, ordinals, admitted, preloaded=frozenset()):
    """Per buffer [(j, phase, score, len)] with the reference's best-of-5 rule."""
    out = []
    for recs, ob in zip(records_per_buffer, ordinals):
        res = []
        for j, w in recs:
            best, bt, bl = -2, 0, 7
            for t, wd in enumerate(w):
                kind, key = wd >> 29, wd & 0xFFFFFF
                if kind == K_NONE:
                    continue
                m = key == 0 or key in preloaded or (key in admitted and admitted[key] < (ob, j, t))
                score, ln = {K_PS: (1000 if m else -1, 7), K_11Z: (1600 if m else 750, 7),
                             K_11I: (1000 if m else -1, 7), K_17: (1800 if m else 1400, 14),
                             K_18: (1800 if m else 1400, 14), K_PL: (1000 if m else -2, 14)}[kind]
                if score > best:
                    best, bt, bl = score, t, ln
            if best >= 0:
                res.append((j, 4 + bt, best, bl))
        out.append(res)
    return out


def resolve_records(records_per_buffer, preloaded=frozenset(), capacity=4096, ordinals=None):
    ordinals = list(range(len(records_per_buffer))) if ordinals is None else ordinals
    admitted = finalize(local_events(records_per_buffer, ordinals), preloaded, capacity)
    return resolve(records_per_buffer, ordinals, admitted, preloaded), set(preloaded) | set(admitted)
