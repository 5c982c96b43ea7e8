/* Minimal stand-in for <GL/glew.h>: the GL stack is not installed in this image and the
 * reference's launcher.h pulls window.h (and thus GL types) in unconditionally.
 * Test infrastructure only - used solely to compile the reference under oracle/_ref. */
#pragma once
typedef unsigned int GLuint; typedef int GLint; typedef unsigned int GLenum; typedef unsigned char GLubyte;
typedef float GLfloat; typedef char GLchar; typedef int GLsizei;
