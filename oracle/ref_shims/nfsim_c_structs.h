/* placeholder, see nfsim_c.h */
