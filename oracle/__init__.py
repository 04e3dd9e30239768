"""TEST INFRASTRUCTURE ONLY: CPU restatement (oracle) of the gauenk/smjp hot path.  See smjp_oracle.py."""
