/* oracle/gibbs_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  (filled in below) */
