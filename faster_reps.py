"""Top-level shim with the reference's script name (`python faster_reps.py in.fasta out.fasta [-d]`,
shannon.py:606)."""
from shannon_b200.faster_reps import find_reps, main  # noqa: F401

if __name__ == '__main__':
    main()
