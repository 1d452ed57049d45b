/* TEST INFRASTRUCTURE (oracle/): forwarding stub so that the reference sources under
 * /root/reference compile for the host.  The TI header of this name is not vendored
 * in the reference; every type it would provide is restated in include/trik_xdm.h. */
#include "trik_xdm.h"
