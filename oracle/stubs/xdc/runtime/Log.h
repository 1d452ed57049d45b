/* TEST INFRASTRUCTURE (oracle/): the reference logs through the XDC runtime
 * (src/vidtranscode_cv.cpp:61,143,...).  On the host these are no-ops. */
#ifndef ORACLE_STUB_XDC_LOG_H_
#define ORACLE_STUB_XDC_LOG_H_
#define Log_error0(fmt)                 ((void)0)
#define Log_error1(fmt, a)              ((void)0)
#define Log_error2(fmt, a, b)           ((void)0)
#define Log_error3(fmt, a, b, c)        ((void)0)
#define Log_error4(fmt, a, b, c, d)     ((void)0)
#define Log_info0(fmt)                  ((void)0)
#define Log_info1(fmt, a)               ((void)0)
#define Log_info2(fmt, a, b)            ((void)0)
#define Log_warning0(fmt)               ((void)0)
#define Log_warning1(fmt, a)            ((void)0)
#endif
