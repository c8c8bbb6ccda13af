/* placeholder, filled in with the native-mode mirror */
#ifndef NATIVE_MIRROR_H
#define NATIVE_MIRROR_H
#endif
