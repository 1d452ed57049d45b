/* TEST INFRASTRUCTURE (oracle/): empty stand-in for <xdc/runtime/Diags.h>. */
