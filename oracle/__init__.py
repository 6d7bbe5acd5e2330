"""TEST INFRASTRUCTURE ONLY: CPU oracle of the Shannon k-mer front end (see shannon_oracle.py).
Nothing under shannon_b200/ may import this package."""
