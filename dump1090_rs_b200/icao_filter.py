"""Mirror of the reference's `icao_filter` module (src/icao_filter.rs)."""
from . import _ffi
from .context import default_context

ICAO_FILTER_ADSB_NT = _ffi.ICAO_FILTER_ADSB_NT   # src/icao_filter.rs:6


def icao_flush(ctx=None):                       # :11-17
    (ctx or default_context()).icao_flush()


def icao_hash(a32: int) -> int:                 # :19-43
    return int(_ffi.lib().b200adsb_icao_hash(a32 & 0xFFFFFFFF))


def icao_filter_add(addr: int, ctx=None):       # :46-62
    (ctx or default_context()).icao_filter_add(addr)


def icao_filter_test(addr: int, ctx=None) -> bool:   # :65-97
    return (ctx or default_context()).icao_filter_test(addr)
