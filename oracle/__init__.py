"""oracle/ — TEST INFRASTRUCTURE ONLY (parity checker + CPU baseline).  See oracle/ngp_oracle.c."""
