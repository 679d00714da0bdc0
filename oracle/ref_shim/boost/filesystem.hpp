// oracle/ref_shim: boost::filesystem is only named by a commented-out block of the reference.
