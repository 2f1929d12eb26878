"""`deflate foo.z` -> `foo`: mirror of pure-zlib's command-line tool (reference: Deflate.hs:15-48 at the
root of the checkout; Haskell port: haskell/app/Deflate.hs).  Drives the incremental decoder exactly as
the reference's `runDecompression` does: feed a strict chunk on NeedMore, append on Chunk, stop on Done
or DecompError.  The chunks are those of `L.readFile` (32 KiB - 16 bytes each,
bytestring's defaultChunkSize) so that the decoder sees what the reference's decoder would see.

    python -m pure_zlib_b200.deflate_cli foo.z
"""
from __future__ import annotations

import sys
from typing import List

from . import zlib as Z

LAZY_CHUNK = 32 * 1024 - 16  # Data.ByteString.Lazy.Internal.defaultChunkSize (32k - 2 words of overhead)


def run_decompression(out, chunks: List[bytes], state, echo=print) -> None:
    """`runDecompression` (Deflate.hs:30-48)."""
    while True:
        if isinstance(state, Z.Done):
            if chunks:
                echo("WARNING: Finished decompression with data left.")
            out.close()
            return
        if isinstance(state, Z.DecompError):
            echo("ERROR: " + str(state.error))
            out.close()
            return
        if isinstance(state, Z.NeedMore):
            if not chunks:
                echo("ERROR: Ran out of data mid-decompression.")
                out.close()
                return
            state = state.feed(chunks.pop(0))
        else:  # Chunk
            out.write(state.data)
            state = state.next()


def main(argv: List[str], echo=print) -> int:
    if len(argv) != 1:
        echo("USAGE: deflate [filename]")
        return 0
    path = argv[0]
    if not path.endswith(".z"):
        echo("Unexpected file name.")
        return 0
    data = open(path, "rb").read()
    chunks = [data[i:i + LAZY_CHUNK] for i in range(0, len(data), LAZY_CHUNK)]
    run_decompression(open(path[:-2], "wb"), chunks, Z.decompress_incremental(), echo)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1:]))
