/* TEST INFRASTRUCTURE ONLY — declarations (no implementations) of the OpenGL / GLU / freeglut entry points that the
 * reference's main.cpp names, so that tests/test_dropin_cpp.py can check that the UNMODIFIED main.cpp compiles
 * against include/dropin/ (this container has no GL headers).  Nothing links against this. */
#ifndef SPHSM_TEST_STUB_FREEGLUT_H
#define SPHSM_TEST_STUB_FREEGLUT_H
typedef unsigned int GLenum, GLbitfield, GLuint;
typedef int GLint, GLsizei;
typedef float GLfloat;
typedef double GLdouble;
typedef unsigned char GLubyte;
enum {
    GL_AMBIENT = 1, GL_BACK, GL_COLOR_BUFFER_BIT, GL_DEPTH_BUFFER_BIT, GL_DEPTH_TEST, GL_DIFFUSE, GL_FILL, GL_FRONT,
    GL_FRONT_AND_BACK, GL_LIGHT0, GL_LIGHTING, GL_LIGHTING_BIT, GL_LIGHT_MODEL_TWO_SIDE, GL_LINES, GL_LINE_LOOP, GL_MODELVIEW,
    GL_POINTS, GL_PROJECTION, GL_RENDERER, GL_SHADING_LANGUAGE_VERSION, GL_SHININESS, GL_SMOOTH, GL_SPECULAR, GL_TRIANGLES,
    GL_VERSION, GLUT_DEPTH, GLUT_DOUBLE, GLUT_DOWN, GLUT_ELAPSED_TIME, GLUT_RGB, GLUT_UP
};
extern "C" {
void glBegin(GLenum); void glEnd(void); void glClear(GLbitfield); void glClearColor(GLfloat, GLfloat, GLfloat, GLfloat);
void glColor3f(GLfloat, GLfloat, GLfloat); void glDisable(GLenum); void glEnable(GLenum); const GLubyte *glGetString(GLenum);
void glLightModelf(GLenum, GLfloat); void glLightfv(GLenum, GLenum, const GLfloat *); void glLineWidth(GLfloat);
void glLoadIdentity(void); void glMaterialf(GLenum, GLenum, GLfloat); void glMaterialfv(GLenum, GLenum, const GLfloat *);
void glMatrixMode(GLenum); void glNormal3f(GLfloat, GLfloat, GLfloat); void glPointSize(GLfloat); void glPolygonMode(GLenum, GLenum);
void glPopAttrib(void); void glPopMatrix(void); void glPushAttrib(GLbitfield); void glPushMatrix(void);
void glRotatef(GLfloat, GLfloat, GLfloat, GLfloat); void glScalef(GLfloat, GLfloat, GLfloat); void glShadeModel(GLenum);
void glTranslatef(GLfloat, GLfloat, GLfloat); void glVertex3f(GLfloat, GLfloat, GLfloat); void glViewport(GLint, GLint, GLsizei, GLsizei);
void gluLookAt(GLdouble, GLdouble, GLdouble, GLdouble, GLdouble, GLdouble, GLdouble, GLdouble, GLdouble);
void gluPerspective(GLdouble, GLdouble, GLdouble, GLdouble);
void glutCloseFunc(void (*)(void)); int glutCreateWindow(const char *); void glutDisplayFunc(void (*)(void)); int glutGet(GLenum);
void glutIdleFunc(void (*)(void)); void glutInit(int *, char **); void glutInitDisplayMode(unsigned int);
void glutInitWindowPosition(int, int); void glutInitWindowSize(int, int); void glutKeyboardFunc(void (*)(unsigned char, int, int));
void glutMainLoop(void); void glutMotionFunc(void (*)(int, int)); void glutMouseFunc(void (*)(int, int, int, int));
void glutPostRedisplay(void); void glutReshapeFunc(void (*)(int, int)); void glutSetWindowTitle(const char *); void glutSwapBuffers(void);
}
#endif
