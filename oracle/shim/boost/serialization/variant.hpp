// oracle shim: serialisation is a no-op under the shim (see pagmo/s11n.hpp); nothing to declare.
