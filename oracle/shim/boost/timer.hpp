// empty stand-in: the reference includes <boost/timer.hpp> but uses nothing from it on the built path
