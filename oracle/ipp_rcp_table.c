#include <xmmintrin.h>
#include <immintrin.h>
#include <stdio.h>
int main(){
  for(int d=1; d<=510; ++d){
    float x=(float)d; float y=x*(1.0f/255.0f);
    float r1=_mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(x)));
    float r2=_mm_cvtss_f32(_mm_rcp_ss(_mm_set_ss(y)));
    printf("%d %.9g %.9g\n", d, r1, r2);
  }
  return 0;
}
