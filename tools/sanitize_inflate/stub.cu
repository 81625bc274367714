#include <stdarg.h>
#include <stdio.h>
void llfe_set_error(const char* fmt, ...) { (void)fmt; }
