"""TEST INFRASTRUCTURE ONLY.  Minimal stand-in for pyfaidx (absent from this image) so that the
unmodified reference's get_fasta_stats (cnvlib/reference.py:859-895) can run on the small synthetic
genomes oracle/make_golden.py writes: Fasta(fname, as_raw=True)[name][start:end] -> str with Python
slice clamping, KeyError for an unknown sequence -- the behaviour of pyfaidx the reference relies on.
The whole file is held in memory; not for real genomes."""


class _Record:
    def __init__(self, name, seq):
        self.name = name
        self._seq = seq

    def __len__(self):
        return len(self._seq)

    def __getitem__(self, key):
        return self._seq[key]


class Fasta:
    def __init__(self, filename, as_raw=False, **kwargs):
        self.filename = filename
        self._records = {}
        name, chunks = None, []
        with open(filename) as handle:
            for line in handle:
                if line.startswith(">"):
                    if name is not None:
                        self._records[name] = "".join(chunks)
                    name, chunks = line[1:].split()[0], []
                else:
                    chunks.append(line.rstrip("\r\n"))
        if name is not None:
            self._records[name] = "".join(chunks)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def __getitem__(self, name):
        try:
            return _Record(name, self._records[name])
        except KeyError:
            raise KeyError(f"{name} not in {self.filename}.") from None

    def keys(self):
        return self._records.keys()
