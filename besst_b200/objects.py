"""Boundary data model: attribute-compatible stand-ins for BESST's
`Contig.contig` (Contig.py:23-38), `Scaffold.scaffold` (Scaffold.py:23-35) and
`Parameter.counters` (Parameter.py:113-124).  When the BESST package is
importable its own classes are used, so objects created here are the very
types the rest of BESST expects; otherwise these slot-identical classes are."""


class contig(object):
    __slots__ = ('name', 'scaffold', 'direction', 'position', 'length',
                 'coverage', 'repeat', 'is_haplotype', 'sequence')

    def __init__(self, contig_name, contig_scaffold=None, contig_direction=None, contig_position=None,
                 contig_length=None, contig_coverage=None, contig_repeat=False, contig_haplotype=False,
                 contig_sequence=None):
        self.name = contig_name
        self.scaffold = contig_scaffold
        self.direction = contig_direction
        self.position = contig_position
        self.length = contig_length
        self.sequence = contig_sequence
        self.coverage = contig_coverage
        self.repeat = contig_repeat
        self.is_haplotype = contig_haplotype


class scaffold(object):
    __slots__ = ('name', 'contigs', 's_length')

    def __init__(self, scaffold_name, scaffold_contigs, scaffold_length):
        self.name = scaffold_name
        self.contigs = scaffold_contigs
        self.s_length = scaffold_length


def classes():
    try:  # pragma: no cover - BESST is not installed in the build image
        from BESST import Contig as _C, Scaffold as _S
        return _C.contig, _S.scaffold
    except Exception:
        return contig, scaffold
