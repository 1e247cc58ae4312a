"""Boundary types: Sequence (reference src/ds/sequence.rs:10-13) and Tile (src/ds/tile.rs:9-11).

Bytes in, bytes out.  Equality and hashing look at `chain` only (sequence.rs:104-117);
results of the aligner carry id=None (sequence.rs:17-19)."""
from typing import Iterable, List, Optional


class Sequence:
    __slots__ = ("chain", "id")

    def __init__(self, chain=b"", id: Optional[str] = None):
        if isinstance(chain, Sequence):
            chain, id = chain.chain, chain.id if id is None else id
        elif isinstance(chain, str):
            chain = chain.encode("ascii")
        self.chain = bytearray(chain)
        self.id = id

    # sequence.rs:21-47
    def push(self, x: int):
        self.chain.append(x)

    def pop(self):
        return self.chain.pop() if self.chain else None

    def extend(self, b: "Sequence"):
        self.chain.extend(b.chain)

    def back(self):
        return self.chain[-1] if self.chain else None

    def __len__(self):
        return len(self.chain)

    def is_empty(self):
        return not self.chain

    def reverse(self):
        self.chain.reverse()

    def __getitem__(self, i):
        return self.chain[i]

    def __iter__(self):
        return iter(self.chain)

    def __eq__(self, other):
        return isinstance(other, Sequence) and self.chain == other.chain

    def __hash__(self):
        return hash(bytes(self.chain))

    def __bytes__(self):
        return bytes(self.chain)

    def __str__(self):
        return self.chain.decode("ascii", "replace")

    def __repr__(self):
        return "Sequence(%r, id=%r)" % (bytes(self.chain), self.id)


class Tile:
    """Vec<Sequence>.  For the batched entry points pair p = (tile[2p], tile[2p+1])
    (the fixture layout of tests/integration.rs:236-242 and examples/from_file.rs:26-27)."""
    __slots__ = ("data",)

    def __init__(self, data: Iterable[Sequence] = ()):
        self.data: List[Sequence] = [s if isinstance(s, Sequence) else Sequence(s) for s in data]

    def push(self, value: Sequence):
        self.data.append(value if isinstance(value, Sequence) else Sequence(value))

    def pop(self):
        return self.data.pop() if self.data else None

    def remove(self, index: int) -> Sequence:
        return self.data.pop(index)

    def size(self):
        return (len(self.data), len(self.data[0]))

    def __len__(self):
        return len(self.data)

    def is_empty(self):
        return not self.data

    def extend(self, b: "Tile"):
        self.data.extend(b.data)

    def __getitem__(self, i):
        return self.data[i]

    def __setitem__(self, i, v):
        self.data[i] = v

    def __iter__(self):
        return iter(self.data)

    def __eq__(self, other):
        return isinstance(other, Tile) and self.data == other.data
