"""FASTA reader with the reference's quirks (ribbit.cpp:269-280): a '>' line closes the previous record
only if its sequence is non-empty, the name is the text between '>' and the first space, all other lines are appended
verbatim (a '\\r' therefore becomes an N base), and after EOF one more record is emitted unconditionally."""


def read_fasta(path):
    names, seqs = [], []
    name, parts = "", []
    with open(path, "rb") as f:
        data = f.read()
    lines = data.split(b"\n")
    if lines and lines[-1] == b"":
        lines.pop()          # getline does not report an empty line after the final newline
    for line in lines:
        if line[:1] == b">":
            seq = b"".join(parts)
            if seq != b"":
                names.append(name); seqs.append(seq)
            sp = line.find(b" ")
            name = (line[1:sp] if sp >= 0 else line[1:]).decode(errors="replace")
            parts = []
        else:
            parts.append(line)
    names.append(name); seqs.append(b"".join(parts))
    return names, seqs
