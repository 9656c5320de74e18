// Headless stand-in for <GL/glew.h>: the reference driver only needs these names to COMPILE;
// its --out path never calls into GL (OpticalFlow.cpp:1072-1074). Test infrastructure only.
#ifndef MOF_SHIM_GLEW_H
#define MOF_SHIM_GLEW_H
typedef unsigned int GLuint;
typedef float GLfloat;
#define GLEW_OK 0
static inline int glewInit(void) { return GLEW_OK; }
#endif
