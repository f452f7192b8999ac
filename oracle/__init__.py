"""CPU oracle — TEST INFRASTRUCTURE ONLY (see vkhrt_oracle.cpp). Never import from vkhrt_b200/."""
