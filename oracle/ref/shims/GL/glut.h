// Headless stand-in for <GL/glut.h>; see glew.h in this directory. Test infrastructure only.
#ifndef MOF_SHIM_GLUT_H
#define MOF_SHIM_GLUT_H
#include <stdio.h>
#include <stdlib.h>
#define GLUT_RGB 0
#define GLUT_DOUBLE 2
#define GLUT_CURSOR_WAIT 7
#define GLUT_CURSOR_INHERIT 100
static inline void glutSetCursor(int) {}
static inline void glutInitDisplayMode(unsigned int) {}
static inline void glutInitWindowSize(int, int) {}
static inline void glutInit(int*, char**) {}
static inline int glutCreateWindow(const char*) { return 0; }
static inline void glutIdleFunc(void (*)(void)) {}
static inline void glutDisplayFunc(void (*)(void)) {}
static inline void glutReshapeFunc(void (*)(int, int)) {}
static inline void glutMouseFunc(void (*)(int, int, int, int)) {}
static inline void glutMotionFunc(void (*)(int, int)) {}
static inline void glutKeyboardFunc(void (*)(unsigned char, int, int)) {}
static inline void glutSpecialFunc(void (*)(int, int, int)) {}
static inline void glutMainLoop(void) { fprintf(stderr, "[ERROR] headless oracle build: no viewer, pass --out\n"); exit(1); }
#endif
