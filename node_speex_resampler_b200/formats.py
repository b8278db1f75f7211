"""Container helpers for the step in front of the resampler. The reference's
fixtures are RIFF/WAVE files named .pcm and its test feeds them header and all (src/test.ts:30-32,
44 header bytes resampled as if they were audio); `wav_pcm` finds the PCM payload instead. Byte
parsing only -- no sample arithmetic happens on the host."""
from __future__ import annotations

import struct
from typing import NamedTuple


class WavPcm(NamedTuple):
    channels: int
    sample_rate: int
    bits_per_sample: int
    format_tag: int      # 1 = integer PCM, 3 = IEEE float, 0xFFFE = extensible
    data: memoryview     # the `data` chunk, interleaved little-endian samples


def wav_pcm(blob) -> WavPcm:
    """The format and the sample bytes of a RIFF/WAVE blob (chunks may come in any order, odd-sized
    chunks are padded to even as the format says; a truncated `data` chunk yields what is there)."""
    view = memoryview(blob).cast("B")
    if len(view) < 12 or bytes(view[0:4]) != b"RIFF" or bytes(view[8:12]) != b"WAVE":
        raise ValueError("not a RIFF/WAVE blob")
    pos, fmt, data = 12, None, None
    while pos + 8 <= len(view):
        tag = bytes(view[pos:pos + 4])
        (size,) = struct.unpack_from("<I", view, pos + 4)
        body = view[pos + 8: pos + 8 + size]
        if tag == b"fmt " and len(body) >= 16:
            fmt = struct.unpack_from("<HHIIHH", body, 0)
        elif tag == b"data":
            data = body
            if fmt is not None:
                break
        pos += 8 + size + (size & 1)
    if fmt is None or data is None:
        raise ValueError("WAVE blob without fmt / data chunk")
    format_tag, channels, rate, _, block_align, bits = fmt
    if block_align:
        data = data[: len(data) - len(data) % block_align]
    return WavPcm(channels, rate, bits, format_tag, data)
